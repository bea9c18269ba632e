"""Uni3DETRTransformer / Uni3DETRTransformerDecoder / UniCrossAtten drop-ins
(reference: projects/mmdet3d_plugin/models/utils/uni3detr_transformer.py) plus the mmcv bricks
the configs name (BaseTransformerLayer, MultiheadAttention, FFN - SURVEY.md A.8).

Parameter names follow SURVEY.md Appendix B so reference checkpoints load. The forward is
restructured for the GPU: the G query groups the reference decodes serially
(uni3detr_transformer.py:115-125) share all weights and never interact, so they are folded
into the batch (sequence s = b*G + g) and every layer runs once over all B*G*nq rows:
sine embedding, self-attention core and the cross-attention sampling block are libu3d_b200
kernels, the plain linear layers are library GEMMs.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..compat import (ATTENTION, FEEDFORWARD_NETWORK, TRANSFORMER, TRANSFORMER_LAYER,
                      TRANSFORMER_LAYER_SEQUENCE, build_from_cfg)


def inverse_sigmoid(x, eps=1e-5):
    """mmdet.models.utils.transformer.inverse_sigmoid."""
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


class MLP(nn.Module):
    """uni3detr_transformer.py:18-30."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))


@ATTENTION.register_module()
class MultiheadAttention(nn.Module):
    """mmcv MultiheadAttention wrapper: parameters live in ``self.attn`` (nn.MultiheadAttention)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout=None,
                 dropout_layer=None, init_cfg=None, batch_first=False, **kwargs):
        super().__init__()
        if dropout is not None:
            attn_drop = dropout
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        if embed_dims // num_heads != 32:
            raise NotImplementedError("u3d_mha_core is specialised for head_dim == 32")
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop)


@FEEDFORWARD_NETWORK.register_module()
class FFN(nn.Module):
    """mmcv FFN: layers = Sequential(Sequential(Linear, act, drop), ..., Linear, drop)."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type="ReLU", inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__()
        assert num_fcs == 2 and act_cfg.get("type", "ReLU") == "ReLU" and add_identity
        self.embed_dims = embed_dims
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True),
                          nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))


@ATTENTION.register_module()
class UniCrossAtten(nn.Module):
    """uni3detr_transformer.py:215-360 - single-point trilinear sample of the voxel volume,
    sigmoid gate, output projection, positional MLP of the (logit) reference point."""

    def __init__(self, embed_dims=256, num_heads=8, num_points=1, num_sweeps=1, cam_sweep_feq=12,
                 voxel_range=(0, 0, 0), im2col_step=64, dropout=0.1, norm_cfg=None, init_cfg=None,
                 batch_first=False, fp16_enabled=False):
        super().__init__()
        if embed_dims % num_heads != 0:
            raise ValueError(f"embed_dims must be divisible by num_heads, but got {embed_dims} "
                             f"and {num_heads}")
        if num_points != 1:
            raise NotImplementedError("num_points != 1 is not used by any Uni3DETR config")
        self.embed_dims, self.num_heads, self.num_points = embed_dims, num_heads, num_points
        self.dropout = nn.Dropout(dropout)
        self.attention_weights = nn.Linear(embed_dims, num_points)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.position_encoder = nn.Sequential(
            nn.Linear(3, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True),
            nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True))
        self.batch_first = batch_first
        if fp16_enabled:
            self.fp16_enabled = fp16_enabled
        self.init_weight()

    def init_weight(self):
        nn.init.constant_(self.attention_weights.weight, 0.)
        nn.init.constant_(self.attention_weights.bias, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.)

    @torch.no_grad()
    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        """Reference layout: query (nq,B,C), value (B,1,C,D,H,W) or (B,C,D,H,W),
        reference_points (B,nq,3) logits. Returns (nq,B,C)."""
        if self.training:
            raise NotImplementedError("UniCrossAtten: autograd is a 'next' row; call .eval()")
        nq, B, C = query.shape
        v = value[:, 0] if value.dim() == 6 else value
        if v.dim() != 5:
            raise NotImplementedError("the 4-D BEV branch is not used by any Uni3DETR config")
        dt = v.dtype
        vol = v.permute(0, 2, 3, 4, 1).contiguous()                       # NDHWC (free if channels_last_3d)
        q = query.permute(1, 0, 2).reshape(B * nq, C).to(dt).contiguous()
        qp = None if query_pos is None else \
            query_pos.permute(1, 0, 2).reshape(B * nq, C).to(dt).contiguous()
        ref = reference_points.reshape(B * nq, 3).float().contiguous()
        s = ops.cross_sample(vol, ref, q, qp, self.attention_weights.weight.float().reshape(-1).contiguous(),
                             float(self.attention_weights.bias.item()), nq)
        out = F.linear(s, self.output_proj.weight.to(dt), self.output_proj.bias.to(dt))
        pe = self.position_encoder
        pf = ref.to(dt)
        for i in (0, 3):
            pf = F.linear(pf, pe[i].weight.to(dt), pe[i].bias.to(dt))
            pf = F.relu(F.layer_norm(pf, (C,), pe[i + 1].weight.to(dt), pe[i + 1].bias.to(dt),
                                     pe[i + 1].eps))
        res = q if residual is None else residual.permute(1, 0, 2).reshape(B * nq, C).to(dt)
        out = out + res + pf
        return out.reshape(B, nq, C).permute(1, 0, 2)


@TRANSFORMER_LAYER.register_module()
class BaseTransformerLayer(nn.Module):
    """mmcv BaseTransformerLayer restricted to the operation order every Uni3DETR config uses:
    ('self_attn','norm','cross_attn','norm','ffn','norm'), post-norm."""

    def __init__(self, attn_cfgs=None, ffn_cfgs=None, operation_order=None,
                 norm_cfg=dict(type="LN"), init_cfg=None, batch_first=False, **kwargs):
        super().__init__()
        expect = ("self_attn", "norm", "cross_attn", "norm", "ffn", "norm")
        if tuple(operation_order) != expect:
            raise NotImplementedError(f"operation_order {operation_order}; supported: {expect}")
        assert norm_cfg.get("type", "LN") == "LN"
        self.operation_order = tuple(operation_order)
        self.attentions = nn.ModuleList([build_from_cfg(dict(c), ATTENTION) for c in attn_cfgs])
        self.embed_dims = self.attentions[0].embed_dims
        ffn = dict(ffn_cfgs)
        ffn.setdefault("type", "FFN")
        ffn.setdefault("embed_dims", self.embed_dims)
        self.ffns = nn.ModuleList([build_from_cfg(ffn, FEEDFORWARD_NETWORK)])
        self.norms = nn.ModuleList([nn.LayerNorm(self.embed_dims) for _ in range(3)])


def get_sine_pos_embed(pos, num_pos_feats=128, temperature=10000):
    """uni3detr_transformer.py:33-65 as differentiable torch ops (the training path; inference uses the
    u3d_sine_embed kernel): pos (..., 3) in [0,1] -> (..., 3*128), [sin v0, cos v1, sin v2, ...] per coordinate."""
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=pos.device)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    v = pos.unsqueeze(-1) * (2 * math.pi) / dim_t                                   # (..., 3, F)
    e = torch.stack((v[..., 0::2].sin(), v[..., 1::2].cos()), dim=-1).flatten(-2)   # interleave
    return e.flatten(-2)


def _wb(lin, dtype):
    return (lin.weight.detach().to(dtype).contiguous(), lin.bias.detach().to(dtype).contiguous())


def _mlp_plan(mlp_layers, dtype):
    return [_wb(l, dtype) for l in mlp_layers]


def _linear_relu(x, w, b):
    """relu(x @ w.T + b) with the bias+ReLU in the GEMM epilogue (cuBLASLt) for 2-D CUDA inputs."""
    if x.is_cuda and x.dim() == 2:
        return torch._addmm_activation(b, x, w.t())
    return torch.relu_(F.linear(x, w, b))


def _run_mlp(x, plan):
    for i, (w, b) in enumerate(plan):
        x = _linear_relu(x, w, b) if i < len(plan) - 1 else F.linear(x, w, b)
    return x


def _ln(a, ln, b=None, c=None, relu=False):
    """act(LayerNorm(a (+b) (+c))): one fused libu3d kernel on CUDA, torch ops otherwise."""
    w, bias, eps = ln
    if a.is_cuda and a.dim() == 2 and a.shape[1] % (256 if a.dtype == torch.bfloat16 else 128) == 0:
        return ops.add_layernorm(a.contiguous(), b, c, w, bias, eps, relu)
    x = a if b is None else a + b
    x = x if c is None else x + c
    y = F.layer_norm(x, (a.shape[-1],), w, bias, eps)
    return torch.relu_(y) if relu else y


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class Uni3DETRTransformerDecoder(nn.Module):
    """uni3detr_transformer.py:133-212."""

    def __init__(self, transformerlayers=None, num_layers=None, return_intermediate=False,
                 init_cfg=None):
        super().__init__()
        if isinstance(transformerlayers, dict):
            transformerlayers = [dict(transformerlayers) for _ in range(num_layers)]
        self.num_layers = num_layers
        self.layers = nn.ModuleList([build_from_cfg(dict(c), TRANSFORMER_LAYER)
                                     for c in transformerlayers])
        self.embed_dims = self.layers[0].embed_dims
        self.return_intermediate = return_intermediate
        self.d_model = d_model = 256
        self.query_scale = MLP(d_model, d_model, d_model, 3)
        self.ref_point_head = MLP(384, d_model, d_model, 3)
        self.compute_dtype = torch.float32
        self.use_tensor_cores = True   # bf16: linears on tcgen05 with fused epilogues (ops.linear_tc)
        self.last_reg_tmp = None       # per-layer raw reg-branch outputs (R, code) f32 of the last forward
        self.last_ref_logits = None    # (R, 3) f32 reference logits each layer started from (+ the final ones)
        self._plan = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.invalidate())

    def invalidate(self):
        self._plan = None

    def train(self, mode=True):
        self._plan = None
        return super().train(mode)

    @torch.no_grad()
    def prepare(self):
        dt = self.compute_dtype
        E = self.embed_dims
        layers = []
        for layer in self.layers:
            mha = layer.attentions[0].attn
            ca = layer.attentions[1]
            w, b = mha.in_proj_weight.detach(), mha.in_proj_bias.detach()
            ffn = layer.ffns[0].layers
            layers.append(dict(
                heads=mha.num_heads,
                in_qk=(w[:2 * E].to(dt).contiguous(), b[:2 * E].to(dt).contiguous()),
                in_v=(w[2 * E:].to(dt).contiguous(), b[2 * E:].to(dt).contiguous()),
                out=_wb(mha.out_proj, dt),
                ln=[(n.weight.detach().to(dt), n.bias.detach().to(dt), n.eps) for n in layer.norms],
                gate_w=ca.attention_weights.weight.detach().float().reshape(-1).contiguous(),
                gate_b=float(ca.attention_weights.bias.detach().float().item()),
                oproj=_wb(ca.output_proj, dt),
                pe=[_wb(ca.position_encoder[0], dt), _wb(ca.position_encoder[3], dt)],
                pe_ln=[(ca.position_encoder[i].weight.detach().to(dt),
                        ca.position_encoder[i].bias.detach().to(dt), ca.position_encoder[i].eps)
                       for i in (1, 4)],
                ffn=[_wb(ffn[0][0], dt), _wb(ffn[1], dt)]))
        self._plan = dict(dtype=dt, layers=layers,
                          query_scale=_mlp_plan(self.query_scale.layers, dt),
                          ref_point_head=_mlp_plan(self.ref_point_head.layers, dt), tc=None)
        dev = self.query_scale.layers[0].weight.device
        if dt == torch.bfloat16 and dev.type == "cuda" and self.use_tensor_cores:
            self._plan["tc"] = self._prepare_tc()
        return self._plan

    def _prepare_tc(self):
        """bf16 serving plan: every nn.Linear as a pre-swizzled tcgen05 operand with its bias (and the
        LayerNorm that follows it) for ops.linear_tc (csrc/linear_tc.cu)."""
        PL = ops.PackedLinear
        E = self.embed_dims
        ln_of = lambda n: (n.weight, n.bias, n.eps)
        layers = []
        for layer in self.layers:
            mha, ca, ffn = layer.attentions[0].attn, layer.attentions[1], layer.ffns[0].layers
            w, b = mha.in_proj_weight.detach(), mha.in_proj_bias.detach()
            pe = ca.position_encoder
            layers.append(dict(
                heads=mha.num_heads,
                in_qk=PL(w[:2 * E], b[:2 * E]), in_v=PL(w[2 * E:], b[2 * E:]),
                out=PL(mha.out_proj.weight, mha.out_proj.bias, ln_of(layer.norms[0])),
                gate_w=ca.attention_weights.weight.detach().float().reshape(-1).contiguous(),
                gate_b=float(ca.attention_weights.bias.detach().float().item()),
                pe0=tuple(t.detach().float().contiguous() for t in (pe[0].weight, pe[0].bias, pe[1].weight, pe[1].bias))
                + (pe[1].eps,),
                pe1=PL(pe[3].weight, pe[3].bias, ln_of(pe[4])),
                oproj=PL(ca.output_proj.weight, ca.output_proj.bias, ln_of(layer.norms[1])),
                ffn0=PL(ffn[0][0].weight, ffn[0][0].bias),
                ffn1=PL(ffn[1].weight, ffn[1].bias, ln_of(layer.norms[2]))))
        mlp = lambda m: [PL(l.weight, l.bias) for l in m.layers]
        return dict(layers=layers, query_scale=mlp(self.query_scale), ref_point_head=mlp(self.ref_point_head))

    @torch.no_grad()
    def forward_batched(self, query, value_ndhwc, reference_points, nq, reg_plans=None):
        """query (B,Q,E), value (B,D,H,W,E) NDHWC, reference_points (B,Q,3) logits with
        Q = G*nq (G independent groups). Returns (L,B,Q,E) states and (L,B,Q,3) logits."""
        if self.training:
            raise NotImplementedError("decoder autograd is a 'next' row; call .eval()")
        p = self._plan
        if p is None or p["dtype"] != self.compute_dtype:
            p = self.prepare()
        dt = p["dtype"]
        B, Q, E = query.shape
        assert Q % nq == 0
        R, n_seq = B * Q, B * (Q // nq)
        out = query.reshape(R, E).to(dt).contiguous()
        ref = reference_points.reshape(R, 3).float().contiguous()
        value = value_ndhwc.to(dt).contiguous()
        self.last_reg_tmp = None
        if p.get("tc") is not None and out.is_cuda and E == 256:
            return self._forward_tc(p["tc"], out, value, ref, B, Q, nq, n_seq, reg_plans)
        inter, inter_ref = [], []
        for lid, L in enumerate(p["layers"]):
            sine = ops.sine_embed(ref, dt)
            qpos = _run_mlp(sine, p["ref_point_head"])
            if lid != 0:
                qpos = _run_mlp(out, p["query_scale"]) * qpos
            # self attention (q = k = x + pos, v = x), post-norm
            qk = F.linear(out + qpos, *L["in_qk"])
            v = F.linear(out, *L["in_v"])
            attn = ops.mha_core(qk[:, :E], qk[:, E:], v, n_seq, nq, L["heads"])
            x = _ln(out, L["ln"][0], F.linear(attn, *L["out"]))
            # cross attention: sample * gate -> proj, + residual + positional MLP
            s = ops.cross_sample(value, ref, x, qpos, L["gate_w"], L["gate_b"], Q)
            o = F.linear(s, *L["oproj"])
            pf = ref.to(dt)
            for (w, b), ln in zip(L["pe"], L["pe_ln"]):
                pf = _ln(F.linear(pf, w, b), ln, relu=True)
            x = _ln(o, L["ln"][1], x, pf)
            # FFN
            h = _linear_relu(x, *L["ffn"][0])
            x = _ln(x, L["ln"][2], F.linear(h, *L["ffn"][1]))
            out = x
            if reg_plans is not None:
                tmp = _run_mlp(out, reg_plans[lid]).float()
                ref = ref + torch.stack((tmp[:, 0], tmp[:, 1], tmp[:, 4]), dim=1)
            if self.return_intermediate:
                inter.append(out.view(B, Q, E))
                inter_ref.append(ref.view(B, Q, 3))
        if self.return_intermediate:
            return torch.stack(inter), torch.stack(inter_ref)
        return out.view(1, B, Q, E), ref.view(1, B, Q, 3)

    def forward_train_batched(self, query, value_ndhwc, reference_points, nq, reg_branches=None):
        """Training-mode decoder under autograd, fp32 (uni3detr_transformer.py:145-212 + mmcv BaseTransformerLayer,
        SURVEY A.8), the G query groups folded into the batch like the inference path. Linears / LayerNorm /
        nn.MultiheadAttention run as torch ops on the module parameters; the UniCrossAtten sampling block is
        autograd.CrossSampleFn (u3d_cross_sample forward, u3d_cross_sample_bwd backward). Dropouts follow the
        modules' probabilities (mmcv: attention dropout + output dropout 0.1, FFN 0.1, UniCrossAtten 0.1)."""
        from .autograd import CrossSampleFn
        B, Q, E = query.shape
        G = Q // nq
        R, n_seq = B * Q, B * G
        out = query.reshape(R, E).float()
        ref = reference_points.reshape(R, 3).float()
        value = value_ndhwc.float().contiguous()
        inter, inter_ref = [], []
        tr = self.training

        def mlp(m, x):
            for i, l in enumerate(m.layers):
                x = F.relu(l(x)) if i < m.num_layers - 1 else l(x)
            return x
        for lid, layer in enumerate(self.layers):
            qpos = mlp(self.ref_point_head, get_sine_pos_embed(ref.sigmoid()))
            if lid != 0:
                qpos = mlp(self.query_scale, out) * qpos
            mha, ca, ffn = layer.attentions[0], layer.attentions[1], layer.ffns[0].layers
            # self attention: nn.MultiheadAttention over (nq, n_seq, E), q = k = x + pos, v = x (+ identity, dropout)
            x3 = out.view(n_seq, nq, E).transpose(0, 1)
            qk = (out + qpos).view(n_seq, nq, E).transpose(0, 1)
            att = mha.attn(qk, qk, value=x3, need_weights=False)[0].transpose(0, 1).reshape(R, E)
            x = layer.norms[0](out + F.dropout(att, mha.attn.dropout, tr))
            # cross attention (uni3detr_transformer.py:318-360)
            s = CrossSampleFn.apply(value, ref, x, qpos, ca.attention_weights.weight, ca.attention_weights.bias, Q)
            o = F.dropout(ca.output_proj(s), ca.dropout.p, tr)
            x = layer.norms[1](o + x + ca.position_encoder(ref))
            # FFN
            h = F.dropout(F.relu(ffn[0][0](x)), ffn[0][2].p, tr)
            x = layer.norms[2](x + F.dropout(ffn[1](h), ffn[2].p, tr))
            out = x
            if reg_branches is not None:
                tmp = reg_branches[lid](out)
                ref = (ref + torch.stack((tmp[:, 0], tmp[:, 1], tmp[:, 4]), dim=1)).detach()
            if self.return_intermediate:
                inter.append(out.view(B, Q, E))
                inter_ref.append(ref.view(B, Q, 3))
        if self.return_intermediate:
            return torch.stack(inter), torch.stack(inter_ref)
        return out.view(1, B, Q, E), ref.view(1, B, Q, 3)

    def uses_tc(self):
        p = self._plan
        if p is None or p["dtype"] != self.compute_dtype:
            p = self.prepare()
        return p.get("tc") is not None

    def reg_plans_of(self, reg_branches):
        """Per-layer reg-branch linears in the form forward_batched consumes: packed tcgen05 operands
        in the bf16 serving mode, (weight, bias) pairs otherwise."""
        lins = [[m for m in br if isinstance(m, nn.Linear)] for br in reg_branches]
        if self.uses_tc():
            return [[ops.PackedLinear(m.weight, m.bias) for m in br] for br in lins]
        return [[_wb(m, self.compute_dtype) for m in br] for br in lins]

    def _forward_tc(self, T, out, value, ref, B, Q, nq, n_seq, reg_plans):
        """bf16 serving path: every linear layer is one ops.linear_tc launch whose epilogue carries the
        bias, ReLU, query_scale multiply, identity adds and LayerNorms that surround it in the reference
        (uni3detr_transformer.py:179-202, :329-360; mmcv BaseTransformerLayer post-norm blocks) - no
        library GEMM and no separate elementwise kernel between the neck and the heads."""
        L_ = ops.linear_tc
        E = self.embed_dims
        R = B * Q
        inter, inter_ref, tmps = [], [], []
        ref_logits = [ref]                       # the reference logits every layer starts from
        rph, qs = T["ref_point_head"], T["query_scale"]
        for lid, L in enumerate(T["layers"]):
            t = L_(L_(ops.sine_embed(ref, torch.bfloat16), rph[0], relu=True), rph[1], relu=True)
            if lid == 0:
                qpos, xq = L_(t, rph[2], add2=out)                       # query_pos, x + query_pos
            else:
                qraw = L_(t, rph[2])
                sc = L_(L_(out, qs[0], relu=True), qs[1], relu=True)
                qpos, xq = L_(sc, qs[2], mul=qraw, add2=out)             # query_scale(x) * raw, x + query_pos
            qk = L_(xq, L["in_qk"])                                      # (R, 2E): q | k
            v = L_(out, L["in_v"])
            attn = ops.mha_core(qk[:, :E], qk[:, E:], v, n_seq, nq, L["heads"])
            x = L_(attn, L["out"], res1=out, ln=True)                    # out-proj + identity + LN
            s = ops.cross_sample(value, ref, x, qpos, L["gate_w"], L["gate_b"], Q)
            pf = L_(ops.pos3_ln_relu(ref, *L["pe0"]), L["pe1"], ln=True, relu_out=True)
            x = L_(s, L["oproj"], res1=x, res2=pf, ln=True)              # output_proj + identity + pos + LN
            out = L_(L_(x, L["ffn0"], relu=True), L["ffn1"], res1=x, ln=True)
            if reg_plans is not None:
                rp = reg_plans[lid]
                tmp, ref = L_(L_(L_(out, rp[0], relu=True), rp[1], relu=True), rp[2], out_f32=True, ref_in=ref)
                tmps.append(tmp)
            ref_logits.append(ref)
            if self.return_intermediate:
                inter.append(out.view(B, Q, E))
                inter_ref.append(ref.view(B, Q, 3))
        self.last_reg_tmp = tmps if reg_plans is not None else None
        self.last_ref_logits = ref_logits
        if self.return_intermediate:
            return torch.stack(inter), torch.stack(inter_ref)
        return out.view(1, B, Q, E), ref.view(1, B, Q, 3)

    def forward(self, query, key, value, query_pos, reference_points=None, reg_branches=None,
                attn_masks=None, **kwargs):
        """Reference layout: query (nq,B,E) seq-first, value (B,1,C,D,H,W), one group."""
        assert query_pos is None
        nq, B, E = query.shape
        v = value[:, 0] if value.dim() == 6 else value
        vol = v.permute(0, 2, 3, 4, 1)
        reg = None
        if reg_branches is not None:
            reg = self.reg_plans_of(reg_branches)
        hs, refs = self.forward_batched(query.permute(1, 0, 2), vol, reference_points, nq, reg)
        return hs.permute(0, 2, 1, 3), refs  # (L,nq,B,E), (L,B,nq,3)


@TRANSFORMER.register_module()
class Uni3DETRTransformer(nn.Module):
    """uni3detr_transformer.py:68-130."""

    def __init__(self, decoder=None, fp16_enabled=False, init_cfg=None, **kwargs):
        super().__init__()
        self.decoder = build_from_cfg(dict(decoder), TRANSFORMER_LAYER_SEQUENCE)
        self.embed_dims = self.decoder.embed_dims
        self.d_model = 256
        if fp16_enabled:
            self.fp16_enabled = fp16_enabled

    def init_weights(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, UniCrossAtten):
                m.init_weight()

    def forward(self, pts_value, query_embed, num_query, reg_branches=None, reg_plans=None,
                **kwargs):
        """pts_value (B,1,C,D,H,W) / (B,C,D,H,W); query_embed (B,Q,256+3).
        Returns inter_states (L,Q,B,256), init_reference (B,Q,3), inter_references (L,B,Q,3)."""
        assert query_embed is not None
        v = pts_value[:, 0] if pts_value.dim() == 6 else pts_value
        vol = v.permute(0, 2, 3, 4, 1)
        reference_points = query_embed[..., self.d_model:]
        query = query_embed[..., :self.d_model]
        if self.training:
            hs, refs = self.decoder.forward_train_batched(query, vol, reference_points, num_query, reg_branches)
            return hs.permute(0, 2, 1, 3), reference_points.float().sigmoid(), refs.sigmoid()
        with torch.no_grad():
            return self._forward_eval(vol, query, reference_points, num_query, reg_branches, reg_plans)

    def _forward_eval(self, vol, query, reference_points, num_query, reg_branches, reg_plans):
        init_reference_out = reference_points.float().sigmoid()
        if reg_plans is None and reg_branches is not None:
            reg_plans = self.decoder.reg_plans_of(reg_branches)
        hs, refs = self.decoder.forward_batched(query, vol, reference_points, num_query, reg_plans)
        return hs.permute(0, 2, 1, 3), init_reference_out, refs.sigmoid()
