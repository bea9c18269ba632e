"""SparseEncoderHD drop-in (reference: projects/mmdet3d_plugin/models/pts_encoder/sparse_encoder_hd.py).

Same registry name, constructor keys and parameter names
(``conv_input.0.weight``, ``encoder_layers.encoder_layer{i}.{j}.conv1.weight`` ...,
SURVEY.md Appendix B; weights in the spconv-1.x layout (kz,ky,kx,Cin,Cout), the 2.x layout
(Cout,kz,ky,kx,Cin) is converted on load). The forward is a different program:

* geometry phase - one rulebook per *resolution* (4 SubM neighbour tables + 3 strided
  rulebooks) instead of one per conv (the reference rebuilds it for each of the 17 SubM
  convs because SparseBasicBlock carries no indice_key);
* feature phase - 21 fused gather-GEMM launches (BN(eval) folded to scale/shift, ReLU and
  the basic-block identity add in the epilogue), no per-layer host sync;
* ``dense()`` writes NDHWC directly (what the dense CNN and the cross-attention read).
"""
import math

import os

import torch
from torch import nn

from .. import ops
from ..compat import MIDDLE_ENCODERS


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return tuple(int(x) for x in v)
    return (int(v),) * 3


class _SparseConvBase(nn.Module):
    subm = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False,
                 indice_key=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _triple(kernel_size)
        self.stride, self.padding = _triple(stride), _triple(padding)
        self.indice_key = indice_key
        if self.kernel_size not in ((3, 3, 3), (1, 1, 1)):
            raise NotImplementedError("sparse conv kernel sizes other than 1 and 3")
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        fan_in = in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        bound = math.sqrt(6.0 / ((1 + 5.0) * fan_in))  # kaiming_uniform(a=sqrt(5))
        nn.init.uniform_(self.weight, -bound, bound)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "weight"
        w = state_dict.get(key)
        if w is not None and tuple(w.shape) != tuple(self.weight.shape) and w.dim() == 5 and \
                w.shape[0] == self.out_channels and w.shape[-1] == self.in_channels:
            state_dict[key] = w.permute(1, 2, 3, 4, 0).contiguous()  # spconv 2.x -> 1.x layout
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class SubMConv3d(_SparseConvBase):
    subm = True


class SparseConv3d(_SparseConvBase):
    subm = False


def make_sparse_convmodule(in_channels, out_channels, kernel_size, indice_key, stride=1,
                           padding=0, conv_type="SubMConv3d", norm_cfg=None,
                           order=("conv", "norm", "act")):
    """mmdet3d.ops.make_sparse_convmodule: Sequential(conv[, BN1d][, ReLU])."""
    norm_cfg = norm_cfg or dict(type="BN1d", eps=1e-3, momentum=0.01)
    layers = []
    for name in order:
        if name == "conv":
            cls = SubMConv3d if conv_type == "SubMConv3d" else SparseConv3d
            layers.append(cls(in_channels, out_channels, kernel_size, stride=stride,
                              padding=padding, bias=False, indice_key=indice_key))
        elif name == "norm":
            layers.append(nn.BatchNorm1d(out_channels, eps=norm_cfg.get("eps", 1e-5),
                                         momentum=norm_cfg.get("momentum", 0.1)))
        elif name == "act":
            layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class SparseBasicBlock(nn.Module):
    """mmdet3d.ops.SparseBasicBlock: conv1-norm1-relu-conv2-norm2-(+identity)-relu, SubM 3^3.

    Upstream inherits mmdet's ``BasicBlock``, which registers its norm layers under the names
    ``build_norm_layer(..., postfix=i)`` returns - ``bn1`` / ``bn2`` for BN - and exposes ``norm1`` /
    ``norm2`` only as properties, so reference checkpoints carry ``...bn1.weight`` etc. Same here;
    state dicts written with the ``norm{1,2}.`` spelling are remapped on load."""

    def __init__(self, inplanes, planes, norm_cfg=None, conv_cfg=None):
        super().__init__()
        norm_cfg = norm_cfg or dict(type="BN1d", eps=1e-3, momentum=0.01)
        eps, mom = norm_cfg.get("eps", 1e-5), norm_cfg.get("momentum", 0.1)
        self.conv1 = SubMConv3d(inplanes, planes, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm1d(planes, eps=eps, momentum=mom)
        self.conv2 = SubMConv3d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm1d(planes, eps=eps, momentum=mom)
        self.relu = nn.ReLU(inplace=True)

    @property
    def norm1(self):
        return self.bn1

    @property
    def norm2(self):
        return self.bn2

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for old, new in (("norm1.", "bn1."), ("norm2.", "bn2.")):
            for key in [k for k in state_dict if k.startswith(prefix + old)]:
                state_dict.setdefault(prefix + new + key[len(prefix + old):], state_dict.pop(key))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


def _fold_bn(bn, conv_bias=None):
    if bn is None:
        return None, (conv_bias.float() if conv_bias is not None else None)
    scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
    shift = bn.bias.float() - bn.running_mean.float() * scale
    if conv_bias is not None:
        shift = shift + conv_bias.float() * scale
    return scale.contiguous(), shift.contiguous()


@MIDDLE_ENCODERS.register_module()
class SparseEncoderHD(nn.Module):
    def __init__(self, in_channels, sparse_shape, order=("conv", "norm", "act"),
                 norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), base_channels=16,
                 output_channels=128,
                 encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                 encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
                 encoder_strides=(2, 2, 2, 1), block_type="conv_module", keep_depth=True,
                 fp16_enabled=False):
        super().__init__()
        assert block_type in ["conv_module", "basicblock"]
        order = tuple(order)
        assert len(order) == 3 and set(order) == {"conv", "norm", "act"}
        if order[0] != "conv":
            raise NotImplementedError("pre-activation order is not used by any Uni3DETR config")
        self.sparse_shape = [int(v) for v in sparse_shape]
        self.in_channels = in_channels
        self.order = order
        self.base_channels = base_channels
        self.output_channels = output_channels
        self.encoder_channels = encoder_channels
        self.encoder_paddings = encoder_paddings
        self.encoder_strides = encoder_strides
        self.stage_num = len(self.encoder_channels)
        self.keep_depth = keep_depth
        if fp16_enabled:
            self.fp16_enabled = fp16_enabled
        self.compute_dtype = torch.float32
        self.use_tensor_cores = True  # bf16 layers with Cin >= 16 run on tcgen05 (u3d_spconv_fwd_packed)

        self.conv_input = make_sparse_convmodule(in_channels, base_channels, 3, norm_cfg=norm_cfg,
                                                 padding=1, indice_key="subm1",
                                                 conv_type="SubMConv3d")
        out_ch = self.make_encoder_layers(make_sparse_convmodule, norm_cfg, base_channels,
                                          block_type=block_type)
        self.conv_out = make_sparse_convmodule(out_ch, output_channels, kernel_size=(1, 1, 1),
                                               stride=(1, 1, 1), norm_cfg=norm_cfg, padding=0,
                                               indice_key="spconv_down2", conv_type="SparseConv3d")
        self._plan = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.invalidate())

    # same construction logic as the reference (sparse_encoder_hd.py:140-214)
    def make_encoder_layers(self, make_block, norm_cfg, in_channels, block_type="conv_module",
                            conv_cfg=dict(type="SubMConv3d")):
        self.encoder_layers = nn.Sequential()
        for i, blocks in enumerate(self.encoder_channels):
            blocks_list = []
            for j, out_channels in enumerate(tuple(blocks)):
                padding = tuple(self.encoder_paddings[i])[j]
                strided = dict(norm_cfg=norm_cfg, stride=self.encoder_strides[i], padding=padding,
                               indice_key=f"spconv{i + 1}", conv_type="SparseConv3d")
                if i != 0 and j == 0 and block_type == "conv_module":
                    blocks_list.append(make_block(in_channels, out_channels, 3, **strided))
                elif block_type == "basicblock":
                    if j == len(blocks) - 1 and i != len(self.encoder_channels) - 1:
                        blocks_list.append(make_block(in_channels, out_channels, 3, **strided))
                    else:
                        blocks_list.append(SparseBasicBlock(out_channels, out_channels,
                                                            norm_cfg=norm_cfg, conv_cfg=conv_cfg))
                else:
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg,
                                                  padding=padding, indice_key=f"subm{i + 1}",
                                                  conv_type="SubMConv3d"))
                in_channels = out_channels
            self.encoder_layers.add_module(f"encoder_layer{i + 1}", nn.Sequential(*blocks_list))
        return out_channels

    # ------------------------------------------------------------------ plan ----
    def invalidate(self):
        self._plan = None

    def train(self, mode=True):
        self.invalidate()
        return super().train(mode)

    def layer_specs(self):
        """Flat list of conv steps: dict(conv, bn, relu, save, add)."""
        specs = []

        def add_module_seq(seq):
            conv = seq[0]
            bn = seq[1] if len(seq) > 1 and isinstance(seq[1], nn.BatchNorm1d) else None
            relu = any(isinstance(m, nn.ReLU) for m in seq)
            specs.append(dict(conv=conv, bn=bn, relu=relu, save=False, add=False))

        add_module_seq(self.conv_input)
        for stage in self.encoder_layers:
            for blk in stage:
                if isinstance(blk, SparseBasicBlock):
                    specs.append(dict(conv=blk.conv1, bn=blk.norm1, relu=True, save=True, add=False))
                    specs.append(dict(conv=blk.conv2, bn=blk.norm2, relu=True, save=False, add=True))
                else:
                    add_module_seq(blk)
        add_module_seq(self.conv_out)
        return specs

    @torch.no_grad()
    def prepare(self, dtype=None):
        dtype = dtype or self.compute_dtype
        steps = []
        for s in self.layer_specs():
            conv = s["conv"]
            k = conv.kernel_size[0] * conv.kernel_size[1] * conv.kernel_size[2]
            w = conv.weight.detach().reshape(k, conv.in_channels, conv.out_channels)
            scale, shift = _fold_bn(s["bn"], conv.bias)
            w = w.to(dtype).contiguous()
            packed, cin = None, conv.in_channels
            if self.use_tensor_cores and dtype == torch.bfloat16 and w.is_cuda:
                if cin < 16 and not steps:
                    # the K-starved stem (Cin = 4/5): zero-pad the reduction dim to 16 so it runs on
                    # the tensor-core kernel too (the input features are padded in forward_voxels)
                    cin = 16
                    w = torch.nn.functional.pad(w, (0, 0, 0, cin - conv.in_channels)).contiguous()
                if ops.spconv_tc_supported(k, cin, conv.out_channels):
                    packed = ops.spconv_pack_weights(w)
                else:
                    cin, w = conv.in_channels, w[:, :conv.in_channels].contiguous()
            x3 = None
            if (self.use_tensor_cores and dtype == torch.float32 and w.is_cuda and
                    os.environ.get("U3D_FP32_STRICT") != "1"):
                # fp32 (BASELINE configs 3 / 5) on the tensor cores: 3xBF16 operand split (ops.PackedConvX3);
                # the stem's 4 / 5 input channels are zero-padded to one 16-wide K block
                wx = w
                if cin < 16:
                    cin = 16
                    wx = torch.nn.functional.pad(w, (0, 0, 0, cin - conv.in_channels)).contiguous()
                if ops.spconv_tc_supported(k, cin, min(conv.out_channels, 128)) and conv.out_channels % 2 == 0 and \
                        (conv.out_channels <= 128 or conv.out_channels % 128 == 0):
                    x3 = ops.PackedConvX3(wx)
                else:
                    cin = conv.in_channels
            steps.append(dict(w=w, packed=packed, x3=x3, scale=scale, shift=shift, relu=s["relu"],
                              save=s["save"], add=s["add"], subm=conv.subm, k=k,
                              cin=cin, cout=conv.out_channels,
                              stride=conv.stride, pad=conv.padding))
        # the 3xBF16 path carries [hi | lo] activation pairs from layer to layer: all layers or none
        use_x3 = all(st["x3"] is not None for st in steps)
        if not use_x3:
            for st in steps:
                st["x3"] = None
        self._plan = dict(dtype=dtype, steps=steps, tc=self.use_tensor_cores, x3=use_x3)
        return self._plan

    # --------------------------------------------------------------- forward ----
    def forward_voxels_train(self, feats, coors, n_rows, cap, vmap, B):
        """Training-mode forward under autograd (sparse_encoder_hd.py:106-138 with train-mode BatchNorm1d over
        the active rows of the whole batch): fp32, exact-size feature matrices (the live row counts
        are read back once per resolution), convs through autograd.SparseConvFn (data gradient = the same
        gather-GEMM over the transposed rulebook, weight gradient = u3d_spconv_wgrad), BN / ReLU / identity add
        as torch ops on the (n, C) matrices, dense() through autograd.ToDenseFn."""
        from .autograd import SparseConvFn, ToDenseFn
        n = int(n_rows)                                           # host sync: once per resolution in training
        x = feats[:n].float().contiguous()
        level = dict(coors=coors, n_t=n_rows, n=n, vmap=vmap, nbr=None, dims=tuple(self.sparse_shape))
        saved = None
        for s in self.layer_specs():
            conv = s["conv"]
            k = conv.kernel_size[0] * conv.kernel_size[1] * conv.kernel_size[2]
            w = conv.weight.reshape(k, conv.in_channels, conv.out_channels)
            if k == 1:
                nbr, out_level = None, level
            elif conv.subm:
                if level["nbr"] is None:
                    level["nbr"] = ops.rulebook_subm(level["coors"], level["n_t"], level["n"], level["vmap"])
                nbr, out_level = level["nbr"], level
            else:
                oc, on, ovm, nbr, ocap = ops.rulebook_down(level["coors"], level["n_t"], level["n"], level["vmap"],
                                                           conv.stride, conv.padding)
                m = int(on)
                nbr = nbr[:, :m].as_subclass(ops.Rulebook) if m < nbr.shape[1] else nbr
                out_level = dict(coors=oc, n_t=on, n=m, vmap=ovm, nbr=None, dims=ovm.dims)
            if s["save"]:
                saved = x
            y = SparseConvFn.apply(x, w, nbr, out_level["n_t"], out_level["n"])
            if conv.bias is not None:
                y = y + conv.bias
            if s["bn"] is not None:
                y = s["bn"](y)                                    # train mode: batch statistics over active rows
            if s["add"]:
                y = y + saved
                saved = None
            x = torch.relu(y) if s["relu"] else y
            level = out_level
        dense = ToDenseFn.apply(x, level["coors"], level["n_t"], level["n"], B, tuple(level["dims"]))
        out = dense.permute(0, 4, 1, 2, 3)
        if not self.keep_depth:
            out = out.sum(dim=2)
        self.last_level = level
        return out

    def forward_voxels(self, feats, coors, n_rows, cap, vmap, B, channels_last=True):
        """feats (cap,Cin) f32/bf16, coors (cap,4) int32, n_rows device int32 (1,), vmap level-0
        VoxelMap. Returns the dense volume (B,C,D,H,W) (channels_last_3d strides by default)."""
        if self.training:
            return self.forward_voxels_train(feats, coors, n_rows, cap, vmap, B)
        with torch.no_grad():
            return self._forward_voxels_eval(feats, coors, n_rows, cap, vmap, B, channels_last)

    def _forward_voxels_eval(self, feats, coors, n_rows, cap, vmap, B, channels_last=True):
        plan = self._plan
        if plan is None or plan["dtype"] != self.compute_dtype or plan["tc"] != self.use_tensor_cores:
            plan = self.prepare()
        dtype = plan["dtype"]
        x = feats.to(dtype).contiguous()
        if plan["steps"][0]["cin"] > x.shape[1]:
            x = torch.nn.functional.pad(x, (0, plan["steps"][0]["cin"] - x.shape[1]))
        x3 = plan.get("x3", False)
        if x3:
            x = ops.split_bf16(x)                # (cap, 2*Cin) bf16 [hi | lo] all the way to dense()
        dims = tuple(self.sparse_shape)
        level = dict(coors=coors, n=n_rows, cap=cap, vmap=vmap, nbr=None, nbr_sorted=None, dims=dims)
        saved = None
        # tile scheduling (csrc/tilesort.cu): the rows-on-N tensor-core kernel takes rulebooks whose 256-slot tiles
        # group rows with similar neighbour masks, so that a tile multiplies (and gathers, and fetches weight images
        # for) few offsets that feed none of its rows: useful slots 6-72 % in natural order, 45-91 % sorted.
        # Round 1 sorted only the thin levels (Cin <= 32): on the 64 / 128-channel levels the lost spatial order of
        # the gathers cost more L2 misses than the padding saved, and a sort cost 0.17 ms per level. Round 2: the
        # sorted table of a SubM level is built straight from the coordinates (ops.rulebook_subm_sorted: no natural
        # table, no permute pass) and the weight-stationary MMA form made the padded units cheaper than the fabric
        # traffic they cause, so every SubM level is sorted (measured per step at batch 32, same box: Cin <= 32 only
        # 15.60 ms; + the 64-channel level 15.47 (64->64 1.12 -> 0.89 ms); + the 128-channel level 15.28 (128->128
        # 0.66 -> 0.57); groups of 8 / 16 scenes 15.3 - 15.4; + the single-use strided tables 15.4 - 15.7).
        # U3D_SORT_TILES=0 keeps the natural order everywhere; U3D_SORT_MAX_CIN / U3D_SORT_DOWN / U3D_SORT_GROUP /
        # U3D_SORT_FUSED tune the policy.
        sort_tiles = (os.environ.get("U3D_SORT_TILES", "1") != "0" and os.environ.get("U3D_TC_KERNEL") != "1")
        sort_max_cin = int(os.environ.get("U3D_SORT_MAX_CIN", "128"))
        sort_down = os.environ.get("U3D_SORT_DOWN", "0") != "0"
        # U3D_SORT_GROUP=g keeps the buckets inside groups of g scenes (measured: 64->64 1.33 -> 1.04 ms per step when
        # that level is sorted in groups of 4, paid back by the +0.29 ms of its sort; neutral on the default policy).
        # Also measured neutral (15.93 vs 15.95 ms per step): building every rulebook / tile sort on a side stream
        # ahead of the convolutions - they depend on coordinates only, but next to a persistent conv grid the
        # table kernels get a sliver of each SM and the convs end up waiting for them.
        sort_group = int(os.environ.get("U3D_SORT_GROUP", "0"))
        sort_fused = os.environ.get("U3D_SORT_FUSED", "1") != "0"   # sorted SubM tables straight from the coordinates
        # U3D_CONV_ZIGZAG=1: every other conv walks its tiles backwards (a layer leaves its LAST rows in L2, the next
        # one would start on them). Measured neutral at batch 32 (64->64: 0.3154 vs 0.3159 ms), so off by default.
        zigzag = os.environ.get("U3D_CONV_ZIGZAG", "0") != "0"
        def geometry_of(st, level):
            """(rulebook, output level) of one conv step; tables are built once per resolution."""
            sortable = (sort_tiles and (st["packed"] is not None or x3) and st["k"] == 27 and st["cout"] <= 128
                        and st["cin"] <= sort_max_cin and (st["subm"] or sort_down))
            if st["k"] == 1:
                nbr, out_level = None, level
                if x3:                           # 1x1x1 conv on the gather kernel: identity table
                    if level.get("ident") is None:
                        level["ident"] = ops.identity_rulebook(level["cap"], x.device)
                    nbr = level["ident"]
            elif st["subm"]:
                if sortable and level["nbr_sorted"] is None:
                    if level["nbr"] is None and sort_fused:
                        # slot-ordered table straight from the coordinates: no natural table to write, re-read, permute
                        level["nbr_sorted"] = ops.rulebook_subm_sorted(level["coors"], level["n"], level["cap"],
                                                                       level["vmap"], sort_group)
                    else:
                        if level["nbr"] is None:
                            level["nbr"] = ops.rulebook_subm(level["coors"], level["n"], level["cap"], level["vmap"])
                        level["nbr_sorted"] = ops.rulebook_sort_tiles(level["nbr"], level["n"], level["cap"],
                                                                      level["coors"], B, sort_group)
                if not sortable and level["nbr"] is None:
                    level["nbr"] = ops.rulebook_subm(level["coors"], level["n"], level["cap"], level["vmap"])
                nbr, out_level = (level["nbr_sorted"] if sortable else level["nbr"]), level
            else:
                oc, on, ovm, nbr, ocap = ops.rulebook_down(level["coors"], level["n"],
                                                           level["cap"], level["vmap"],
                                                           st["stride"], st["pad"],
                                                           sorted_group=sort_group if (sortable and sort_fused) else None)
                if sortable and not sort_fused:
                    nbr = ops.rulebook_sort_tiles(nbr, on, ocap, oc, B, sort_group)
                out_level = dict(coors=oc, n=on, cap=ocap, vmap=ovm, nbr=None, nbr_sorted=None, dims=ovm.dims)
            return nbr, out_level

        # Geometry first: every rulebook / tile sort / strided coordinate set depends on the voxel COORDINATES only,
        # so the whole chain (4 SubM tables, 3 strided tables, the sorts) is issued before the first convolution.
        # These are small-CTA, low-register kernels that share the SMs with the two FPS launches of the detector's
        # side streams; the persistent conv kernels (one 55 K-register CTA per SM) cannot, and used to stall behind
        # the FPS blocks when they were interleaved with the table builds (U3D_GEOMETRY_FIRST=0: interleaved).
        geom = None
        if os.environ.get("U3D_GEOMETRY_FIRST", "1") != "0":
            geom, lv = [], level
            for st in plan["steps"]:
                g = geometry_of(st, lv)
                geom.append(g)
                lv = g[1]
        rev = False
        for i, st in enumerate(plan["steps"]):
            nbr, out_level = geom[i] if geom is not None else geometry_of(st, level)
            if st["save"]:
                saved = x
            res = saved if st["add"] else None
            if x3:
                rev = zigzag and not rev
                x = ops.spconv_fwd_packed_x3(x, nbr, out_level["n"], out_level["cap"], st["x3"], st["scale"],
                                             st["shift"], residual=res, relu=st["relu"], reverse=rev)
            elif st["packed"] is not None:
                rev = zigzag and not rev
                x = ops.spconv_fwd_packed(x, nbr, out_level["n"], out_level["cap"], st["packed"],
                                          st["k"], st["cin"], st["cout"], st["scale"], st["shift"],
                                          residual=res, relu=st["relu"], reverse=rev)
            else:
                x = ops.spconv_fwd(x, nbr, out_level["n"], out_level["cap"], st["w"], st["scale"],
                                   st["shift"], residual=res, relu=st["relu"])
            if st["add"]:
                saved = None
            level = out_level
        if x3:
            x = ops.merge_bf16(x)                                  # back to fp32 rows
        dense = ops.sparse_to_dense(x, level["coors"], level["n"], level["cap"], B, level["dims"],
                                    channels_last=True)          # (B,D,H,W,C)
        out = dense.permute(0, 4, 1, 2, 3)                       # logical (B,C,D,H,W)
        if not channels_last:
            out = out.contiguous()
        if not self.keep_depth:
            out = out.sum(dim=2)
        self.last_level = level
        return out

    def forward(self, voxel_features, coors, batch_size):
        """Reference API (sparse_encoder_hd.py:106-138): exact-size (N,C) features and (N,4)
        int32 [b,z,y,x] coordinates -> (B,C,D,H,W)."""
        coors = coors.int().contiguous()
        n = coors.shape[0]
        B = int(batch_size)
        n_rows = torch.tensor([n], dtype=torch.int32, device=coors.device)
        vmap = ops.voxmap_build(coors, n_rows, n, B, self.sparse_shape)
        return self.forward_voxels(voxel_features, coors, n_rows, max(n, 1), vmap, B)
