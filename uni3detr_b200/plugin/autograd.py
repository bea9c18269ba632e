"""torch.autograd bindings of the libu3d_b200 kernels that have a hand-written backward.

Reference: in the reference these gradients come from torch.autograd through spconv's `indice_conv`
(models/pts_encoder/sparse_encoder_hd.py:106-138), `SparseConvTensor.dense()` (:133) and `F.grid_sample`
(models/utils/uni3detr_transformer.py:345). Here:
  * SparseConvFn  - forward u3d_spconv_fwd; data gradient = the SAME gather-GEMM kernel over the transposed
                    rulebook (u3d_rulebook_transpose) with W_k^T; weight gradient = u3d_spconv_wgrad;
  * ToDenseFn     - forward u3d_sparse_to_dense (NDHWC scatter), backward = gather of the volume gradient;
  * CrossSampleFn - forward u3d_cross_sample (trilinear gather x sigmoid gate), backward
                    u3d_cross_sample_bwd (8-corner scatter-add, gate and reference-point gradients).
fp32 only: the training step keeps these ops in fp32 (the dense CNN and the linears go through cuDNN /
cuBLAS autograd).
"""
import torch

from .. import ops


class SparseConvFn(torch.autograd.Function):
    """y = sum_k x[nbr[k]] @ w[k]  (no bias / norm / activation: those stay in autograd-visible torch ops).
    x (n_in, Cin) f32, w (K, Cin, Cout) f32, nbr (K, n_out) int32 rulebook or None (1x1x1 conv)."""

    @staticmethod
    def forward(ctx, x, w, nbr, n_out_t, n_out):
        x = x.contiguous()
        w = w.contiguous()
        y = ops.spconv_fwd(x, nbr, n_out_t, n_out, w)
        ctx.save_for_backward(x, w)
        ctx.nbr, ctx.n_out_t, ctx.n_out = nbr, n_out_t, n_out
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        nbr, n_out_t, n_out = ctx.nbr, ctx.n_out_t, ctx.n_out
        dy = dy.contiguous()
        n_in = x.shape[0]
        K, Cin, Cout = w.shape
        dx = dw = None
        if nbr is None:                                            # pointwise: plain matrix products
            if ctx.needs_input_grad[0]:
                dx = dy @ w[0].t()
            if ctx.needs_input_grad[1]:
                dw = (x.t() @ dy).unsqueeze(0)
            return dx, dw, None, None, None
        if ctx.needs_input_grad[0]:
            n_in_t = torch.tensor([n_in], dtype=torch.int32, device=x.device)
            nbr_t = ops.rulebook_transpose(nbr, n_out_t, n_out, n_in)
            dx = ops.spconv_fwd(dy, nbr_t, n_in_t, n_in, w.transpose(1, 2).contiguous())
        if ctx.needs_input_grad[1]:
            dw = ops.spconv_wgrad(x, dy, nbr, n_out_t, n_out, K, Cin, Cout)
        return dx, dw, None, None, None


class ToDenseFn(torch.autograd.Function):
    """(n, C) rows at coors (n,4)[b,z,y,x] -> dense (B, D, H, W, C)."""

    @staticmethod
    def forward(ctx, feats, coors, n_t, n, B, dims):
        out = ops.sparse_to_dense(feats.contiguous(), coors, n_t, n, B, dims, channels_last=True)
        ctx.save_for_backward(coors)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, d_out):
        (coors,) = ctx.saved_tensors
        c = coors[:ctx.n].long()
        return d_out[c[:, 0], c[:, 1], c[:, 2], c[:, 3]], None, None, None, None, None


class CrossSampleFn(torch.autograd.Function):
    """out = trilinear_sample(value, sigmoid(ref)) * sigmoid((query + query_pos) . gate_w + gate_b)."""

    @staticmethod
    def forward(ctx, value_ndhwc, ref, query, query_pos, gate_w, gate_b, Q):
        value_ndhwc, ref, query = value_ndhwc.contiguous(), ref.contiguous(), query.contiguous()
        query_pos = None if query_pos is None else query_pos.contiguous()
        gw = gate_w.reshape(-1).contiguous()
        out = ops.cross_sample(value_ndhwc, ref, query, query_pos, gw, float(gate_b.item()), Q)
        ctx.save_for_backward(value_ndhwc, ref, query, query_pos if query_pos is not None else query.new_empty(0), gw, gate_b)
        ctx.has_pos, ctx.Q, ctx.gw_shape = query_pos is not None, Q, gate_w.shape
        return out

    @staticmethod
    def backward(ctx, d_out):
        value, ref, query, qpos, gw, gb = ctx.saved_tensors
        qpos = qpos if ctx.has_pos else None
        d_value, d_q, d_gw, d_gb, d_ref = ops.cross_sample_bwd(value, ref, query, qpos, gw, float(gb.item()), ctx.Q,
                                                               d_out.contiguous(), ctx.needs_input_grad[0])
        return (d_value, d_ref if ctx.needs_input_grad[1] else None, d_q, d_q if ctx.has_pos else None,
                d_gw.reshape(ctx.gw_shape), d_gb.reshape(gb.shape), None)
