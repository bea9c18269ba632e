"""SECOND3D / SECOND3DFPN drop-ins (reference: models/backbones/second_3d.py,
models/necks/second3d_fpn.py).

The dense 3-D CNN between the sparse encoder and the decoder stays on the library path
(cuDNN through torch), as SURVEY.md §0.4 / §8a rows a9-a10 prescribe; what changes is how it
is driven: eval-mode BatchNorm3d folded into the conv weights, bf16 (or fp32) NDHWC
(`channels_last_3d`) activations end to end so no layout transposes are inserted, and the
(1,3,3) convs of SECOND3D issued as NHWC 2-D convs over (B*D) slices (same memory, no copy).
Parameter names match the reference (``blocks.{i}.{0,3,..}.weight``, ``deblocks.{i}.0.weight``,
``extra_blocks.{0,3,6}.weight``).
"""
import numpy as np
import os

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..compat import BACKBONES, NECKS


FUSE_CONV_BIAS_RELU = True


def _norm(norm_cfg, ch):
    t = norm_cfg.get("type", "BN3d")
    if t not in ("BN3d", "BN", "naiveSyncBN3d"):
        raise NotImplementedError(f"norm type {t}")
    return nn.BatchNorm3d(ch, eps=norm_cfg.get("eps", 1e-5), momentum=norm_cfg.get("momentum", 0.1))


def _fold(conv_w, bn, transposed=False, conv_bias=None):
    scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
    shift = bn.bias.float() - bn.running_mean.float() * scale
    if conv_bias is not None:      # conv_cfg / upsample_cfg / extra_conv with bias=True
        shift = shift + conv_bias.float() * scale
    w = conv_w.float()
    w = w * (scale.view(1, -1, 1, 1, 1) if transposed else scale.view(-1, 1, 1, 1, 1))
    return w, shift


class _FoldedConv:
    """conv(+folded BN)+ReLU on channels_last_3d tensors."""

    def __init__(self, conv, bn, dtype, as2d):
        transposed = isinstance(conv, nn.ConvTranspose3d)
        w, b = _fold(conv.weight.detach(), bn, transposed,
                     conv.bias.detach() if conv.bias is not None else None)
        self.transposed = transposed
        self.stride, self.padding = conv.stride, conv.padding
        self.b = b.to(dtype).contiguous()
        self.b32 = b.float().contiguous()
        k = conv.kernel_size
        self.as2d = bool(as2d and k[0] == 1 and conv.stride[0] == 1 and conv.padding[0] == 0)
        if self.as2d:
            self.w = w[:, :, 0].to(dtype).contiguous(memory_format=torch.channels_last)
        else:
            self.w = w.to(dtype).contiguous(memory_format=torch.channels_last_3d)
        # fp32 (BASELINE configs 3 / 5, 1e-3 parity): "3xTF32" - w = w_hi + w_lo, x = x_hi + x_lo with the hi
        # parts exactly representable in TF32; conv(x_hi,w_hi) + conv(x_lo,w_hi) + conv(x_hi,w_lo) on the tensor
        # cores with fp32 accumulation drops only the lo*lo term (2^-22 relative): fp32-grade results at
        # tensor-core speed instead of cuDNN's FFMA kernels. U3D_FP32_STRICT=1 keeps the plain fp32 convs.
        self.w_hi = self.w_lo = None
        if dtype == torch.float32 and self.w.is_cuda and os.environ.get("U3D_FP32_STRICT") != "1":
            from .. import ops
            self.w_hi, self.w_lo = ops.split_tf32(self.w)

    def __call__(self, x):
        if self.w.dtype == torch.float32:
            if self.w_hi is not None and x.is_cuda:
                return torch.relu_(self._conv3x(x, self.b))
            # strict fp32: keep cuDNN off its TF32 tensor-core path
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self._run(x)
        return self._run(x)

    def raw(self, x):
        """Transposed conv WITHOUT its bias / ReLU epilogue (both are applied by the fused level merge,
        ops.bias_act_sum): one cuDNN dgrad launch instead of dgrad + bias-add + clamp."""
        assert self.transposed
        if self.w.dtype == torch.float32:
            if self.w_hi is not None and x.is_cuda:
                return self._conv3x(x, None)
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self._raw(x)
        return self._raw(x)

    def _conv(self, x, w, b):
        if self.as2d:
            B, C, D, H, W = x.shape
            x2 = x.permute(0, 2, 1, 3, 4).reshape(B * D, C, H, W)  # view on NDHWC memory
            if self.transposed:
                y = F.conv_transpose2d(x2, w, b, stride=self.stride[1:])
            else:
                y = F.conv2d(x2, w, b, stride=self.stride[1:], padding=self.padding[1:])
            return y.reshape(B, D, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
        if self.transposed:
            return F.conv_transpose3d(x, w, b, stride=self.stride)
        return F.conv3d(x, w, b, stride=self.stride, padding=self.padding)

    def _conv3x(self, x, b):
        from .. import ops
        x = x if ops.is_dense(x) else x.contiguous(memory_format=torch.channels_last_3d)
        x_hi, x_lo = ops.split_tf32(x)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=True):
            y = self._conv(x_hi, self.w_hi, b)
            y += self._conv(x_lo, self.w_hi, None)
            y += self._conv(x_hi, self.w_lo, None)
        return y

    def _raw(self, x):
        if self.as2d:
            B, C, D, H, W = x.shape
            x2 = x.permute(0, 2, 1, 3, 4).reshape(B * D, C, H, W)
            y = F.conv_transpose2d(x2, self.w, None, stride=self.stride[1:])
            return y.reshape(B, D, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
        return F.conv_transpose3d(x, self.w, None, stride=self.stride)

    def _run(self, x):
        # conv + bias + ReLU as ONE cuDNN fused op (torch.cudnn_convolution_relu) where it exists
        # (forward convs on CUDA); transposed convs keep the bias/ReLU as two elementwise launches.
        fused = FUSE_CONV_BIAS_RELU and x.is_cuda and not self.transposed
        if self.as2d:
            B, C, D, H, W = x.shape
            x2 = x.permute(0, 2, 1, 3, 4).reshape(B * D, C, H, W)  # view on NDHWC memory
            if self.transposed:
                y = torch.relu_(F.conv_transpose2d(x2, self.w, self.b, stride=self.stride[1:]))
            elif fused:
                y = torch.cudnn_convolution_relu(x2, self.w, self.b, self.stride[1:], self.padding[1:], (1, 1), 1)
            else:
                y = torch.relu_(F.conv2d(x2, self.w, self.b, stride=self.stride[1:], padding=self.padding[1:]))
            return y.reshape(B, D, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
        if self.transposed:
            return torch.relu_(F.conv_transpose3d(x, self.w, self.b, stride=self.stride))
        if fused:
            return torch.cudnn_convolution_relu(x, self.w, self.b, self.stride, self.padding, (1, 1, 1), 1)
        return torch.relu_(F.conv3d(x, self.w, self.b, stride=self.stride, padding=self.padding))


def _fold_sequential(seq, dtype, as2d):
    mods = list(seq)
    out, i = [], 0
    while i < len(mods):
        conv = mods[i]
        assert isinstance(conv, (nn.Conv3d, nn.ConvTranspose3d)), type(conv)
        assert isinstance(mods[i + 1], nn.BatchNorm3d) and isinstance(mods[i + 2], nn.ReLU)
        out.append(_FoldedConv(conv, mods[i + 1], dtype, as2d))
        i += 3
    return out


class _PlanMixin:
    compute_dtype = torch.float32
    conv2d_trick = True

    def invalidate(self):
        self._plan = None

    def train(self, mode=True):
        self._plan = None
        return super().train(mode)


@BACKBONES.register_module()
class SECOND3D(_PlanMixin, nn.Module):
    def __init__(self, in_channels=128, out_channels=[128, 128, 256], layer_nums=[3, 5, 5],
                 layer_strides=[2, 2, 2], is_cascade=True,
                 norm_cfg=dict(type="BN3d", eps=1e-3, momentum=0.01),
                 conv_cfg=dict(type="Conv3d", bias=False), init_cfg=None, pretrained=None):
        super().__init__()
        assert len(layer_strides) == len(layer_nums) == len(out_channels)
        conv_cfg = dict(conv_cfg)
        self.kernel_type = conv_cfg.get("type", "Conv3d")
        if self.kernel_type != "Conv3d":
            raise NotImplementedError("SECOND3D: only conv_cfg.type='Conv3d' (all shipped configs)")
        kernel = tuple(conv_cfg.pop("kernel", (1, 3, 3)))
        bias = conv_cfg.get("bias", False)
        in_filters = list(in_channels) if isinstance(in_channels, (list, tuple)) \
            else [in_channels, *out_channels[:-1]]
        padding = tuple((k - 1) // 2 for k in kernel)
        self.is_cascade = is_cascade
        blocks = []
        for i, layer_num in enumerate(layer_nums):
            block = [nn.Conv3d(in_filters[i], out_channels[i], kernel,
                               stride=(1, layer_strides[i], layer_strides[i]), padding=padding,
                               bias=bias),
                     _norm(norm_cfg, out_channels[i]), nn.ReLU(inplace=True)]
            for _ in range(layer_num):
                block += [nn.Conv3d(out_channels[i], out_channels[i], kernel, padding=padding,
                                    bias=bias),
                          _norm(norm_cfg, out_channels[i]), nn.ReLU(inplace=True)]
            blocks.append(nn.Sequential(*block))
        self.blocks = nn.ModuleList(blocks)
        self._plan = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.invalidate())
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    @torch.no_grad()
    def prepare(self):
        self._plan = dict(dtype=self.compute_dtype, as2d=self.conv2d_trick,
                          blocks=[_fold_sequential(b, self.compute_dtype, self.conv2d_trick)
                                  for b in self.blocks])
        return self._plan

    def forward(self, x):
        if self.training:
            # second_3d.py:89-114 under autograd: the nn.Conv3d / BatchNorm3d(train) / ReLU stacks on cuDNN
            outs = []
            for blk in self.blocks:
                y = blk(x)
                outs.append(y)
                if self.is_cascade:
                    x = y
            return tuple(outs)
        with torch.no_grad():
            return self._forward_eval(x)

    def _forward_eval(self, x):
        p = self._plan
        if p is None or p["dtype"] != self.compute_dtype or p["as2d"] != self.conv2d_trick:
            p = self.prepare()
        x = x.to(self.compute_dtype).contiguous(memory_format=torch.channels_last_3d)
        outs = []
        for blk in p["blocks"]:
            y = x
            for conv in blk:
                y = conv(y)
            outs.append(y)
            if self.is_cascade:
                x = y
        return tuple(outs)


@NECKS.register_module()
class SECOND3DFPN(_PlanMixin, nn.Module):
    def __init__(self, in_channels=[128, 128, 256], out_channels=[256, 256, 256],
                 upsample_strides=[1, 2, 4], norm_cfg=dict(type="BN3d", eps=1e-3, momentum=0.01),
                 upsample_cfg=dict(type="deconv3d", bias=False),
                 conv_cfg=dict(type="Conv3d", bias=False), extra_conv=None,
                 use_conv_for_no_stride=False, use_for_distill=False, init_cfg=None):
        super().__init__()
        assert len(out_channels) == len(upsample_strides) == len(in_channels)
        if "3d" not in upsample_cfg.get("type", "deconv3d") or "3d" not in conv_cfg.get("type", "Conv3d"):
            raise NotImplementedError("SECOND3DFPN: only the 3-D layer types of the shipped configs")
        if use_for_distill:
            raise NotImplementedError("use_for_distill belongs to the OV-Uni3DETR family (out of scope)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.fp16_enabled = False
        deblocks = []
        for i, oc in enumerate(out_channels):
            stride = upsample_strides[i]
            if stride > 1 or (stride == 1 and not use_conv_for_no_stride):
                layer = nn.ConvTranspose3d(in_channels[i], oc, (1, stride, stride),
                                           stride=(1, stride, stride),
                                           bias=upsample_cfg.get("bias", False))
            else:
                s = int(np.round(1 / stride))
                layer = nn.Conv3d(in_channels[i], oc, (1, s, s), stride=(1, s, s),
                                  bias=conv_cfg.get("bias", False))
            deblocks.append(nn.Sequential(layer, _norm(norm_cfg, oc), nn.ReLU(inplace=True)))
        self.deblocks = nn.ModuleList(deblocks)
        self.extra_conv = dict(extra_conv) if extra_conv is not None else None
        if self.extra_conv is not None:
            ec = dict(self.extra_conv)
            self.layer_num = ec.pop("num_conv")
            kernel = tuple(ec.pop("kernel", (3, 3, 3)))
            if "sep_kernel" in ec:
                raise NotImplementedError("sep_kernel is not used by any Uni3DETR config")
            padding = tuple((k - 1) // 2 for k in kernel)
            blocks = []
            for _ in range(self.layer_num):
                blocks += [nn.Conv3d(out_channels[-1], out_channels[-1], kernel, padding=padding,
                                     bias=ec.get("bias", False)),
                           _norm(norm_cfg, out_channels[-1]), nn.ReLU(inplace=True)]
            self.extra_blocks = nn.Sequential(*blocks)
        self._plan = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.invalidate())
        for m in self.modules():
            if isinstance(m, (nn.Conv3d, nn.ConvTranspose3d)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    @torch.no_grad()
    def prepare(self):
        dt, a2 = self.compute_dtype, self.conv2d_trick
        self._plan = dict(dtype=dt, as2d=a2,
                          deblocks=[_fold_sequential(d, dt, a2)[0] for d in self.deblocks],
                          extra=_fold_sequential(self.extra_blocks, dt, a2)
                          if self.extra_conv is not None else [])
        return self._plan

    def forward(self, x):
        assert len(x) == len(self.in_channels)
        if self.training:
            # second3d_fpn.py:112-143 under autograd: upsample every level, sum, extra 3x3x3 convs
            ups = [d(xi) for d, xi in zip(self.deblocks, x)]
            out = ups[0]
            for u in ups[1:]:
                out = out + u
            if self.extra_conv is not None:
                out = self.extra_blocks(out)
            return out
        with torch.no_grad():
            return self._forward_eval(x)

    def _forward_eval(self, x):
        p = self._plan
        if p is None or p["dtype"] != self.compute_dtype or p["as2d"] != self.conv2d_trick:
            p = self.prepare()
        xs = [xi.to(self.compute_dtype).contiguous(memory_format=torch.channels_last_3d) for xi in x]
        if xs[0].is_cuda and len(xs) <= 3 and self.out_channels[0] % 8 == 0:
            # level merge as ONE libu3d kernel over NDHWC rows: sum_i act_i(up_i + bias_i); the transposed
            # convs run bare (cuDNN dgrad only), their folded-BN shift and ReLU happen in the merge
            ups, biases, relus = [], [], []
            for d, xi in zip(p["deblocks"], xs):
                if d.transposed:
                    ups.append(d.raw(xi).permute(0, 2, 3, 4, 1).contiguous())
                    biases.append(d.b32)
                    relus.append(True)
                else:
                    ups.append(d(xi).permute(0, 2, 3, 4, 1).contiguous())
                    biases.append(None)
                    relus.append(False)
            out = ops.bias_act_sum(ups, biases, relus).permute(0, 4, 1, 2, 3)
        else:
            ups = [d(xi) for d, xi in zip(p["deblocks"], xs)]
            out = ups[0].permute(0, 2, 3, 4, 1)
            for u in ups[1:]:
                out = out + u.permute(0, 2, 3, 4, 1)
            out = out.permute(0, 4, 1, 2, 3)
        for conv in p["extra"]:
            out = conv(out)
        return out
