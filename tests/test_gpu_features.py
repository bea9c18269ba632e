"""GPU parity, floating-point half of the path: sparse conv (+BN/ReLU/residual epilogue), dense(),
sine embedding, self-attention core, UniCrossAtten sampling, decoder/head, dense CNN.
CUDA (C ABI) vs oracle/model.py and vs the golden vectors produced by the reference's own files.
Tolerances (BASELINE.json north_star): 1e-3 relative fp32, 1e-2 bf16 - measured here as
max|a-b| / max|b| per tensor, the number each assert states."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gen_weights import gen_state_dict
from oracle import geometry as G
from oracle import model as M

pytestmark = pytest.mark.gpu
DEV = "cuda"
T = torch.from_numpy


def relerr(a, b):
    a = a.detach().float().cpu()
    b = torch.as_tensor(b).detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def rand_coors(n, dims, B, seed):
    rng = np.random.default_rng(seed)
    D, H, W = dims
    lin = np.sort(rng.choice(B * D * H * W, size=n, replace=False))
    return np.stack([lin // (D * H * W), (lin // (H * W)) % D, (lin // W) % H, lin % W], 1).astype(np.int32)


def conv_case(n, dims, B, cin, cout, seed, K=27):
    g = torch.Generator().manual_seed(seed)
    coors = rand_coors(n, dims, B, seed)
    x = torch.randn(n, cin, generator=g)
    w = torch.randn(3 if K == 27 else 1, 3 if K == 27 else 1, 3 if K == 27 else 1, cin, cout, generator=g) / (cin * K / 4) ** 0.5
    scale = 0.5 + torch.rand(cout, generator=g)
    shift = 0.1 * torch.randn(cout, generator=g)
    return coors, x, w, scale, shift


def oracle_conv(x, nbr, w, n_out, scale, shift, residual, relu):
    y = M.sparse_conv(x, nbr, w, n_out) * scale + shift
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


TC_KERNELS = ("0", "1")   # U3D_TC_KERNEL: 0 = rows-on-N (spconv_tn.cu) where it applies, 1 = rows-on-M (spconv_tc.cu)


class tc_kernel:
    """Select the tcgen05 sparse-conv kernel variant for the calls inside the block."""
    def __init__(self, which):
        self.which = which

    def __enter__(self):
        import os
        self.prev = os.environ.get("U3D_TC_KERNEL")
        os.environ["U3D_TC_KERNEL"] = self.which

    def __exit__(self, *a):
        import os
        if self.prev is None:
            os.environ.pop("U3D_TC_KERNEL", None)
        else:
            os.environ["U3D_TC_KERNEL"] = self.prev


CONV_SHAPES = [(4, 16), (5, 16), (16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128)]


@pytest.mark.parametrize("cin,cout", CONV_SHAPES)
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
def test_spconv_subm(cin, cout, dtype, tol):
    from uni3detr_b200 import ops
    dims, B, n = (8, 24, 24), 2, 3000
    coors, x, w, scale, shift = conv_case(n, dims, B, cin, cout, cin * 1000 + cout)
    if dtype == torch.bfloat16:                      # same rounded operands on both sides
        x, w = x.bfloat16().float(), w.bfloat16().float()
    nbr_ref = G.subm_rulebook(coors, dims)
    c = T(coors).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, n, B, dims)
    nbr = ops.rulebook_subm(c, n_rows, n, vm)
    res = torch.randn(n, cout, generator=torch.Generator().manual_seed(1))
    if dtype == torch.bfloat16:
        res = res.bfloat16().float()
    for residual, relu in [(None, True), (res, True), (None, False)]:
        ref = oracle_conv(x, nbr_ref, w, n, scale, shift, residual, relu)
        for impl in (1, 0):
            y = ops.spconv_fwd(x.to(DEV, dtype), nbr, n_rows, n, w.reshape(27, cin, cout).to(DEV, dtype).contiguous(),
                               scale.to(DEV), shift.to(DEV),
                               residual=None if residual is None else residual.to(DEV, dtype), relu=relu, impl=impl)
            assert relerr(y, ref) < tol, (impl, residual is not None, relu, relerr(y, ref))
        if dtype == torch.bfloat16 and ops.spconv_tc_supported(27, cin, cout):     # tcgen05 kernel
            wp = ops.spconv_pack_weights(w.reshape(27, cin, cout).to(DEV, dtype).contiguous())
            for kern in TC_KERNELS:
                with tc_kernel(kern):
                    y = ops.spconv_fwd_packed(x.to(DEV, dtype), nbr, n_rows, n, wp, 27, cin, cout, scale.to(DEV),
                                              shift.to(DEV),
                                              residual=None if residual is None else residual.to(DEV, dtype),
                                              relu=relu)
                assert relerr(y, ref) < tol, ("tc", kern, residual is not None, relu, relerr(y, ref))


@pytest.mark.parametrize("cin,cout", [(16, 32), (32, 64), (64, 128)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
def test_spconv_down(cin, cout, dtype, tol):
    from uni3detr_b200 import ops
    dims, B, n = (9, 24, 26), 2, 2500
    coors, x, w, scale, shift = conv_case(n, dims, B, cin, cout, 7 + cin)
    if dtype == torch.bfloat16:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    c = T(coors).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, n, B, dims)
    oc, n_out, ovm, nbr, ocap = ops.rulebook_down(c, n_rows, n, vm, (2, 2, 2), (0, 1, 1))
    roc, rnbr, rod = G.down_rulebook(coors, dims, (2, 2, 2), (0, 1, 1))
    m = len(roc)
    ref = oracle_conv(x, rnbr, w, m, scale, shift, None, True)
    for impl in (1, 0):
        y = ops.spconv_fwd(x.to(DEV, dtype), nbr, n_out, ocap, w.reshape(27, cin, cout).to(DEV, dtype).contiguous(),
                           scale.to(DEV), shift.to(DEV), relu=True, impl=impl)
        assert relerr(y[:m], ref) < tol, (impl, relerr(y[:m], ref))
    if dtype == torch.bfloat16:
        wp = ops.spconv_pack_weights(w.reshape(27, cin, cout).to(DEV, dtype).contiguous())
        for kern in TC_KERNELS:
            with tc_kernel(kern):
                y = ops.spconv_fwd_packed(x.to(DEV, dtype), nbr, n_out, ocap, wp, 27, cin, cout, scale.to(DEV),
                                          shift.to(DEV), relu=True)
            assert relerr(y[:m], ref) < tol, ("tc", kern, relerr(y[:m], ref))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
def test_spconv_pointwise_and_dense(dtype, tol):
    """conv_out (1x1x1, 128->256) + SparseConvTensor.dense() in both layouts."""
    from uni3detr_b200 import ops
    dims, B, n, cin, cout = (5, 12, 10), 2, 700, 128, 256
    coors, x, w, scale, shift = conv_case(n, dims, B, cin, cout, 3, K=1)
    if dtype == torch.bfloat16:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    ref = F.relu((x @ w.reshape(cin, cout)) * scale + shift)
    y = ops.spconv_fwd(x.to(DEV, dtype), None, n_rows, n, w.reshape(1, cin, cout).to(DEV, dtype).contiguous(),
                       scale.to(DEV), shift.to(DEV), relu=True)
    assert relerr(y, ref) < tol
    if dtype == torch.bfloat16:
        wp = ops.spconv_pack_weights(w.reshape(1, cin, cout).to(DEV, dtype).contiguous())
        y2 = ops.spconv_fwd_packed(x.to(DEV, dtype), None, n_rows, n, wp, 1, cin, cout, scale.to(DEV), shift.to(DEV),
                                   relu=True)
        assert relerr(y2, ref) < tol
    c = T(coors).to(DEV)
    dref = torch.zeros(B, *dims, cout)
    cl = T(coors.astype(np.int64))
    dref[cl[:, 0], cl[:, 1], cl[:, 2], cl[:, 3]] = y.float().cpu()
    d1 = ops.sparse_to_dense(y, c, n_rows, n, B, dims, channels_last=True)
    d2 = ops.sparse_to_dense(y, c, n_rows, n, B, dims, channels_last=False)
    torch.testing.assert_close(d1.float().cpu(), dref, rtol=0, atol=0)
    torch.testing.assert_close(d2.float().cpu(), dref.permute(0, 4, 1, 2, 3), rtol=0, atol=0)


def test_spconv_live_count_smaller_than_capacity():
    """Rows beyond the device-side live count must not be touched (fixed-capacity buffers)."""
    from uni3detr_b200 import ops
    dims, B, n, cap = (6, 10, 10), 1, 200, 333
    coors, x, w, scale, shift = conv_case(n, dims, B, 16, 16, 5)
    c = torch.cat([T(coors), torch.zeros(cap - n, 4, dtype=torch.int32)]).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, cap, B, dims)
    nbr = ops.rulebook_subm(c, n_rows, cap, vm)
    xin = torch.cat([x, torch.full((cap - n, 16), float("nan"))]).to(DEV)
    out = torch.full((cap, 16), -5.0, device=DEV)
    ops.spconv_fwd(xin, nbr, n_rows, cap, w.reshape(27, 16, 16).to(DEV).contiguous(), scale.to(DEV), shift.to(DEV),
                   relu=True, out=out)
    ref = oracle_conv(x, G.subm_rulebook(coors, dims), w, n, scale, shift, None, True)
    assert relerr(out[:n], ref) < 1e-4
    assert bool((out[n:] == -5.0).all())


def test_sine_embed_golden(golden):
    from uni3detr_b200 import ops
    pos = T(golden["sine_in"]).double()
    ref_logit = torch.log(pos / (1 - pos)).float().to(DEV)
    out = ops.sine_embed(ref_logit.reshape(-1, 3).contiguous())
    np.testing.assert_allclose(out.cpu().numpy().reshape(golden["sine_out"].shape), golden["sine_out"], atol=2e-5)


# bf16 kernels: "v2" = mha_tc2.cu (default: persistent, P in TMEM, TMA loads); "v1q0" / "v1q1" = the first tcgen05 kernel
# (U3D_MHA_V1=1) with one CTA looping over the query blocks / one CTA per block; fp32 = the SIMT kernel
@pytest.mark.parametrize("variant", ["v2", "v1q0", "v1q1"])
@pytest.mark.parametrize("seq_len,n_seq", [(300, 8), (900, 2), (5, 4), (64, 3), (129, 1), (300, 40), (161, 5), (1024, 1)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
def test_mha_core(seq_len, n_seq, dtype, tol, variant, monkeypatch):
    from uni3detr_b200 import ops
    if variant != "v2":
        if dtype == torch.float32:
            pytest.skip("fp32 has one kernel")
        monkeypatch.setenv("U3D_MHA_V1", "1")
        monkeypatch.setenv("U3D_MHA_QSPLIT", variant[-1])
    g = torch.Generator().manual_seed(seq_len)
    heads, E = 8, 256
    qk = torch.randn(n_seq * seq_len, 2 * E, generator=g).to(dtype)
    v = torch.randn(n_seq * seq_len, E, generator=g).to(dtype)
    out = ops.mha_core(qk.to(DEV)[:, :E], qk.to(DEV)[:, E:], v.to(DEV), n_seq, seq_len, heads)

    def split(t):
        return t.float().view(n_seq, seq_len, heads, 32).permute(0, 2, 1, 3)
    ref = F.scaled_dot_product_attention(split(qk[:, :E]), split(qk[:, E:]), split(v))
    ref = ref.permute(0, 2, 1, 3).reshape(n_seq * seq_len, E)
    assert relerr(out, ref) < tol


def ca_state_dict():
    from test_oracle_golden import ca_state_dict as f
    return f()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1.5e-2)])
def test_uni_cross_atten_golden(golden, dtype, tol):
    """Product UniCrossAtten (cross_sample kernel + proj) vs the output of the reference's own class."""
    from uni3detr_b200.plugin.transformer import UniCrossAtten
    m = UniCrossAtten(embed_dims=256, num_heads=8, num_points=1, dropout=0.1).eval()
    m.load_state_dict(ca_state_dict())
    m = m.to(DEV)
    out = m(T(golden["ca_query"]).to(DEV, dtype), None, T(golden["ca_value"]).to(DEV, dtype),
            query_pos=T(golden["ca_qpos"]).to(DEV, dtype), reference_points=T(golden["ca_ref"]).to(DEV))
    assert relerr(out, golden["ca_out"]) < tol


def test_cross_sample_out_of_bounds_corners():
    """grid_sample zeros padding at the volume border (align_corners=False)."""
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, D, H, W, C, Q = 2, 3, 4, 5, 256, 64
    vol = torch.randn(B, D, H, W, C, generator=g)
    ref = torch.randn(B * Q, 3, generator=g) * 6          # many saturate to the borders
    q = torch.randn(B * Q, C, generator=g)
    gw, gb = torch.randn(C, generator=g) * 0.05, 0.1
    out = ops.cross_sample(vol.to(DEV), ref.to(DEV), q.to(DEV), None, gw.to(DEV), gb, Q)
    grid = ((ref.sigmoid() - 0.5) * 2).view(B, 1, 1, Q, 3)
    samp = F.grid_sample(vol.permute(0, 4, 1, 2, 3), grid, align_corners=False)[:, :, 0, 0].permute(0, 2, 1)
    gate = (q @ gw + gb).sigmoid().view(B, Q, 1)
    assert relerr(out.view(B, Q, C), samp * gate) < 1e-4


def head_fixture(golden):
    from test_oracle_golden import golden_head_cfg
    from uni3detr_b200 import compat, register_all
    register_all()
    cfg0, sd, (nq, ncls, code, L) = golden_head_cfg(golden)
    layer = dict(type="BaseTransformerLayer",
                 attn_cfgs=[dict(type="MultiheadAttention", embed_dims=256, num_heads=8, dropout=0.1),
                            dict(type="UniCrossAtten", num_points=1, embed_dims=256, num_sweeps=1)],
                 ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=64, num_fcs=2, ffn_drop=0.1,
                               act_cfg=dict(type="ReLU", inplace=True)),
                 norm_cfg=dict(type="LN"),
                 operation_order=("self_attn", "norm", "cross_attn", "norm", "ffn", "norm"))
    pc = [-3.2, -0.2, -2., 3.2, 6.2, 0.56]
    hcfg = dict(type="Uni3DETRHead", num_query=nq, num_classes=ncls, in_channels=256, code_size=code,
                with_box_refine=True, as_two_stage=False, sync_cls_avg_factor=True,
                transformer=dict(type="Uni3DETRTransformer", fp16_enabled=False,
                                 decoder=dict(type="Uni3DETRTransformerDecoder", num_layers=L,
                                              return_intermediate=True, transformerlayers=layer)),
                bbox_coder=dict(type="NMSFreeCoder", pc_range=pc, post_center_range=pc, max_num=12, alpha=0.2,
                                num_classes=ncls),
                loss_cls=dict(type="SoftFocalLoss", use_sigmoid=True))
    head = compat.build_from_cfg(hcfg, compat.HEADS).eval()
    sd = dict(sd)
    sd["code_weights"] = head.code_weights.data
    head.load_state_dict(sd, strict=True)
    return head.to(DEV)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 3e-2)])
def test_head_forward_golden(golden, dtype, tol):
    """Product Uni3DETRHead/Transformer/Decoder vs outputs of the reference's own classes."""
    head = head_fixture(golden)
    head.set_compute_dtype(dtype)
    feats = T(golden["head_feats"]).to(DEV, dtype).contiguous(memory_format=torch.channels_last_3d)
    outs = head(feats, None, T(golden["head_fps"]).to(DEV), random_point=T(golden["head_rand"]).to(DEV))
    for k in ("all_cls_scores", "all_bbox_preds", "all_iou_preds"):
        assert outs[k].shape == golden["head_" + k].shape
        assert relerr(outs[k], golden["head_" + k]) < tol, (k, relerr(outs[k], golden["head_" + k]))


def test_nms_free_coder_golden(golden):
    head = head_fixture(golden)
    outs = {k: T(golden["head_" + k]).to(DEV) for k in ("all_cls_scores", "all_bbox_preds", "all_iou_preds")}
    res = head.bbox_coder.decode(outs)
    for i, r in enumerate(res):
        np.testing.assert_array_equal(r["labels"].cpu().numpy(), golden[f"coder_{i}_labels"])
        for k in ("bboxes", "scores", "ious"):
            np.testing.assert_allclose(r[k].cpu().numpy(), golden[f"coder_{i}_{k}"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 2e-2)])
def test_dense_cnn_golden(golden, dtype, tol):
    """Product SECOND3D/SECOND3DFPN (folded BN, NDHWC, 2-D conv trick) vs the reference's classes."""
    from test_oracle_golden import DENSE_BCFG, DENSE_NCFG, dense_state_dicts
    from uni3detr_b200 import compat, register_all
    register_all()
    bsd, nsd = dense_state_dicts()
    bb = compat.build_from_cfg(dict(type="SECOND3D", **DENSE_BCFG), compat.BACKBONES).eval()
    nk = compat.build_from_cfg(dict(type="SECOND3DFPN", **DENSE_NCFG), compat.NECKS).eval()
    bb.load_state_dict(bsd, strict=True)
    nk.load_state_dict(nsd, strict=True)
    bb, nk = bb.to(DEV), nk.to(DEV)
    bb.compute_dtype = nk.compute_dtype = dtype
    xs = bb(T(golden["dense_in"]).to(DEV))
    for i, x in enumerate(xs):
        assert relerr(x, golden[f"dense_bb{i}"]) < tol
    y = nk(xs)
    assert relerr(y, golden["dense_out"]) < tol


def test_decode_fixed_equals_decode(golden):
    head = head_fixture(golden)
    outs = {k: T(golden["head_" + k]).to(DEV) for k in ("all_cls_scores", "all_bbox_preds", "all_iou_preds")}
    boxes, scores, labels, mask = head.bbox_coder.decode_fixed(outs)
    for i, r in enumerate(head.bbox_coder.decode(outs)):
        torch.testing.assert_close(boxes[i][mask[i]], r["bboxes"])
        torch.testing.assert_close(scores[i][mask[i]], r["scores"])
        assert bool((labels[i][mask[i]] == r["labels"]).all())


@pytest.mark.parametrize("kern", TC_KERNELS)
@pytest.mark.parametrize("cin,cout,n", [(256, 256, 700), (256, 512, 300), (128, 256, 129), (64, 64, 1), (16, 16, 128),
                                        (128, 64, 257), (256, 128, 385), (32, 16, 255), (64, 32, 513)])
def test_spconv_tc_wide_and_ragged(cin, cout, n, kern):
    """tcgen05 kernels: Cin blocks > 1, Cout = 512 (two UMMA N halves), Cout < Cin, partial last tile (128- and
    256-row tiles), live count < capacity."""
    from uni3detr_b200 import ops
    dims, B, cap = (6, 12, 12), 1, n + 77
    coors, x, w, scale, shift = conv_case(n, dims, B, cin, cout, n)
    x, w = x.bfloat16().float(), w.bfloat16().float()
    c = torch.cat([T(coors), torch.zeros(cap - n, 4, dtype=torch.int32)]).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, cap, B, dims)
    nbr = ops.rulebook_subm(c, n_rows, cap, vm)
    wp = ops.spconv_pack_weights(w.reshape(27, cin, cout).to(DEV, torch.bfloat16).contiguous())
    xin = torch.cat([x, torch.full((cap - n, cin), float("nan"))]).to(DEV, torch.bfloat16)
    out = torch.full((cap, cout), -5.0, device=DEV, dtype=torch.bfloat16)
    with tc_kernel(kern):
        ops.spconv_fwd_packed(xin, nbr, n_rows, cap, wp, 27, cin, cout, scale.to(DEV), shift.to(DEV), relu=False,
                              out=out)
    ref = oracle_conv(x, G.subm_rulebook(coors, dims), w, n, scale, shift, None, False)
    assert relerr(out[:n], ref) < 1e-2, relerr(out[:n], ref)
    assert bool((out[n:] == -5.0).all())


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 32), (64, 64), (128, 128)])
def test_spconv_tc_many_tiles_per_cta(cin, cout):
    """Persistent loop: more 256-row tiles than SMs (every CTA runs several tiles, both accumulator
    buffers, ring wrap-around, rulebook-slice double buffer), residual + ReLU; the two tcgen05 kernels
    must agree with the fp32-accumulate SIMT kernel on the same bf16 operands."""
    from uni3detr_b200 import ops
    dims, B, n = (24, 64, 64), 2, 100_000
    coors, x, w, scale, shift = conv_case(n, dims, B, cin, cout, 11 + cin)
    c = T(coors).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, n, B, dims)
    nbr = ops.rulebook_subm(c, n_rows, n, vm)
    xb = x.to(DEV, torch.bfloat16)
    wb = w.reshape(27, cin, cout).to(DEV, torch.bfloat16).contiguous()
    res = torch.randn(n, cout, generator=torch.Generator().manual_seed(2)).to(DEV, torch.bfloat16)
    ref = ops.spconv_fwd(xb, nbr, n_rows, n, wb, scale.to(DEV), shift.to(DEV), residual=res, relu=True, impl=1)
    wp = ops.spconv_pack_weights(wb)
    ys = {}
    for kern in TC_KERNELS:
        with tc_kernel(kern):
            y = ops.spconv_fwd_packed(xb, nbr, n_rows, n, wp, 27, cin, cout, scale.to(DEV), shift.to(DEV),
                                      residual=res, relu=True)
        assert relerr(y, ref) < 1e-2, (kern, relerr(y, ref))
        ys[kern] = y
    # sorted tiles (csrc/tilesort.cu): a scheduling permutation only - every row must come out IDENTICAL to
    # the natural-order run of the same kernel (value equality: the sign of an exact zero may differ)
    srt = ops.rulebook_sort_tiles(nbr, n_rows, n)
    with tc_kernel("0"):
        y = ops.spconv_fwd_packed(xb, srt, n_rows, n, wp, 27, cin, cout, scale.to(DEV), shift.to(DEV),
                                  residual=res, relu=True)
    assert bool((y.float() == ys["0"].float()).all())


def _tile_key(mask):
    key = np.zeros_like(mask)
    for line in range(9):
        key |= (((mask >> (3 * line)) & 7) != 0).astype(mask.dtype) << line
    for c in range(3):
        key |= ((mask & (0x1249249 << c)) != 0).astype(mask.dtype) << (9 + c)
    return key


@pytest.mark.parametrize("n,cap,dims,B", [(5000, 5077, (8, 24, 24), 2), (300, 300, (6, 10, 10), 1), (1, 130, (4, 4, 4), 1),
                                          (60_000, 60_000, (24, 64, 64), 2)])
def test_rulebook_sort_tiles(n, cap, dims, B):
    """Tile scheduling: slot_row is a permutation of the live rows ordered by the 12-bit neighbour signature,
    the sorted table is the natural table gathered through it, tile masks are the OR over 128 slots.
    Integer work: exact."""
    from uni3detr_b200 import ops
    coors = rand_coors(n, dims, B, n + 1)
    c = torch.cat([T(coors), torch.zeros(cap - n, 4, dtype=torch.int32)]).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, cap, B, dims)
    nbr = ops.rulebook_subm(c, n_rows, cap, vm)
    srt = ops.rulebook_sort_tiles(nbr, n_rows, cap)
    nat = nbr[:, :n].cpu().numpy()
    np.testing.assert_array_equal(nat, G.subm_rulebook(coors, dims))          # natural table untouched
    slot_row = srt.slot_row[:n].cpu().numpy()
    np.testing.assert_array_equal(np.sort(slot_row), np.arange(n))            # permutation of the live rows
    mask = ((nat >= 0).astype(np.int64) << np.arange(27)[:, None]).sum(0)
    key = _tile_key(mask)[slot_row]
    assert bool((np.diff(key) >= 0).all())                                     # bucketed by signature
    got = srt[:, :n].cpu().numpy()
    np.testing.assert_array_equal(got, nat[:, slot_row])
    nt = (n + 127) // 128
    act = np.zeros((27, nt * 128), bool)
    act[:, :n] = got >= 0
    want = (act.reshape(27, nt, 128).any(2).astype(np.int64) << np.arange(27)[:, None]).sum(0)
    tm = srt.tile_mask[:nt].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    np.testing.assert_array_equal(tm, want)


def _random_boxes(n, n_cls, seed):
    rng = np.random.default_rng(seed)
    centres = rng.random((12, 2)) * 8                      # clustered -> many overlaps
    which = rng.integers(0, 12, n)
    b = np.zeros((n, 7), np.float32)
    b[:, :2] = centres[which] + rng.normal(0, 0.35, (n, 2))
    b[:, 2] = rng.random(n)
    b[:, 3:6] = 0.6 + rng.random((n, 3)) * 1.4
    b[:, 6] = rng.uniform(-np.pi, np.pi, n)
    return b, rng.random(n).astype(np.float32), rng.integers(0, n_cls, n).astype(np.int64)


@pytest.mark.parametrize("n,n_cls,thr", [(200, 3, 0.5), (65, 1, 0.2), (1, 2, 0.5), (300, 10, 0.5)])
def test_nms3d_bev_vs_oracle(n, n_cls, thr):
    """u3d_nms3d_bev (all classes, all scenes, one launch pair) == per-class oracle nms3d."""
    from oracle import postproc as PP
    from uni3detr_b200 import ops
    scenes = [_random_boxes(n, n_cls, 10 * n + s) for s in range(2)]
    bs, ls, vs, refs = [], [], [], []
    for b, s, l in scenes:
        valid = np.ones(n, bool)
        valid[n // 7::9] = False                             # some rows are not candidates
        order = np.lexsort((-s, np.where(valid, l, n_cls)))  # (label asc, score desc), invalid last
        b, s, l, valid = b[order], s[order], l[order], valid[order]
        keep = np.zeros(n, bool)
        for j in range(n_cls):
            idx = np.nonzero((l == j) & valid)[0]
            if len(idx):
                keep[idx[PP.nms3d(b[idx], s[idx], thr)]] = True
        bs.append(b); ls.append(l); vs.append(valid); refs.append(keep)
    keep = ops.nms3d_bev(T(np.stack(bs)).to(DEV), T(np.stack(ls)).int().to(DEV), T(np.stack(vs)).to(DEV), thr)
    np.testing.assert_array_equal(keep.cpu().numpy(), np.stack(refs))


def test_get_bboxes_nms_vs_oracle(golden):
    """Uni3DETRHead.get_bboxes with post_processing 'nms' (+score_thr, num_thr) vs the oracle's
    restatement of uni3detr_head.py:827-918, on the reference-generated head outputs."""
    from oracle import postproc as PP
    head = head_fixture(golden)
    outs = {k: T(golden["head_" + k]).to(DEV) for k in ("all_cls_scores", "all_bbox_preds", "all_iou_preds")}
    # make the decoded boxes overlap: shrink the centre spread, enlarge the sizes
    outs["all_bbox_preds"] = outs["all_bbox_preds"].clone()
    outs["all_bbox_preds"][..., [0, 1]] *= 0.3
    outs["all_bbox_preds"][..., [2, 3, 5]] = outs["all_bbox_preds"][..., [2, 3, 5]].clamp(-1, 1) + 0.5
    for pp in (dict(type="nms", nms_thr=0.3), dict(type="nms", nms_thr=0.1, score_thr=0.3, num_thr=4),
               dict(type="nms", nms_thr=0.5, score_thr=[0.2, 0.4, 0.1])):
        head.post_processing = pp
        got = head.get_bboxes(outs, [{}, {}])
        dec = head.bbox_coder.decode(outs)
        for i, (gb, gs, gl) in enumerate(got):
            d = {k: v.cpu().numpy() for k, v in dec[i].items()}
            rb, rs, rl = PP.get_bboxes_nms(d, head.num_classes, pp)
            assert len(gs) == len(rs), (pp, len(gs), len(rs))
            np.testing.assert_array_equal(gl.cpu().numpy(), rl)
            np.testing.assert_allclose(gs.cpu().numpy(), rs, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(gb.cpu().numpy(), rb, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_add_layernorm(dtype, tol):
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(3)
    a, b, c = (torch.randn(1000, 256, generator=g).to(dtype) for _ in range(3))
    w, bias = (1 + 0.1 * torch.randn(256, generator=g)).to(dtype), (0.1 * torch.randn(256, generator=g)).to(dtype)
    for res, relu in [((None, None), False), ((b, None), False), ((b, c), False), ((None, None), True)]:
        x = a.float()
        for t in res:
            if t is not None:
                x = x + t.float()
        ref = F.layer_norm(x, (256,), w.float(), bias.float(), 1e-5)
        ref = F.relu(ref) if relu else ref
        out = ops.add_layernorm(a.to(DEV), None if res[0] is None else res[0].to(DEV),
                                None if res[1] is None else res[1].to(DEV), w.to(DEV), bias.to(DEV), 1e-5, relu)
        assert relerr(out, ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
def test_bias_act_sum(dtype, tol):
    """Fused SECOND3DFPN level merge: sum_i act_i(x_i + b_i) vs torch (fp32 math, one rounding)."""
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(9)
    shape, C = (2, 3, 5, 7, 256), 256
    xs = [torch.randn(shape, generator=g).to(dtype) for _ in range(3)]
    bs = [None, torch.randn(C, generator=g), torch.randn(C, generator=g)]
    relus = [False, True, True]
    ref = torch.zeros(shape)
    for x, b, r in zip(xs, bs, relus):
        t = x.float() + (b if b is not None else 0.0)
        ref = ref + (F.relu(t) if r else t)
    y = ops.bias_act_sum([x.to(DEV) for x in xs], [None if b is None else b.to(DEV) for b in bs], relus)
    assert relerr(y, ref) < tol
    y1 = ops.bias_act_sum([xs[0].to(DEV)], [bs[1].to(DEV)], [True])          # single operand
    assert relerr(y1, F.relu(xs[0].float() + bs[1])) < tol


@pytest.mark.parametrize("B,per_group", [(4, 1), (5, 2), (3, 8)])
def test_rulebook_sort_tiles_grouped(B, per_group):
    """Signature buckets kept inside groups of consecutive scenes: slot_row permutes every
    group's row range onto itself, keys are sorted inside a group, table / masks as in the global variant."""
    from uni3detr_b200 import ops
    dims, n = (8, 24, 24), 9000
    coors = rand_coors(n, dims, B, 77)
    c = T(coors).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, n, B, dims)
    nbr = ops.rulebook_subm(c, n_rows, n, vm)
    srt = ops.rulebook_sort_tiles(nbr, n_rows, n, c, B, per_group)
    nat = nbr[:, :n].cpu().numpy()
    slot_row = srt.slot_row[:n].cpu().numpy()
    np.testing.assert_array_equal(np.sort(slot_row), np.arange(n))
    grp = coors[:, 0] // per_group
    np.testing.assert_array_equal(grp[slot_row], grp)                  # a group's rows stay in its range
    mask = ((nat >= 0).astype(np.int64) << np.arange(27)[:, None]).sum(0)
    key = grp[slot_row].astype(np.int64) * 4096 + _tile_key(mask)[slot_row]
    assert bool((np.diff(key) >= 0).all())
    np.testing.assert_array_equal(srt[:, :n].cpu().numpy(), nat[:, slot_row])


@pytest.mark.parametrize("n,cap,dims,B,per_group", [(5000, 5077, (8, 24, 24), 2, 0), (300, 300, (6, 10, 10), 1, 0),
                                                    (1, 130, (4, 4, 4), 1, 0), (60_000, 60_000, (24, 64, 64), 2, 0),
                                                    (9000, 9000, (8, 24, 24), 5, 2), (9000, 9100, (8, 24, 24), 3, 1)])
def test_rulebook_subm_sorted_direct(n, cap, dims, B, per_group):
    """The tile-sorted SubM table built straight from coordinates + VoxelMap (no natural table): slot_row is a permutation
    bucketed by (scene group, signature), the table is the ORACLE's natural table gathered through it, tile masks are
    the OR over 128 slots. Rows keep the reference's first-appearance order (perm != identity). Integer work: exact."""
    from uni3detr_b200 import ops
    coors = rand_coors(n, dims, B, n + 3)
    if per_group:                                     # grouped buckets need scene-major rows
        coors = coors[np.argsort(coors[:, 0], kind="stable")]
    c = torch.cat([T(coors), torch.zeros(cap - n, 4, dtype=torch.int32)]).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, cap, B, dims)
    srt = ops.rulebook_subm_sorted(c, n_rows, cap, vm, per_group)
    nat = G.subm_rulebook(coors, dims)
    slot_row = srt.slot_row[:n].cpu().numpy()
    np.testing.assert_array_equal(np.sort(slot_row), np.arange(n))
    mask = ((nat >= 0).astype(np.int64) << np.arange(27)[:, None]).sum(0)
    grp = coors[:, 0] // per_group if per_group else np.zeros(n, np.int64)
    if per_group:
        np.testing.assert_array_equal(grp[slot_row], grp)
    key = grp[slot_row].astype(np.int64) * 4096 + _tile_key(mask)[slot_row]
    assert bool((np.diff(key) >= 0).all())
    got = srt[:, :n].cpu().numpy()
    np.testing.assert_array_equal(got, nat[:, slot_row])
    nt = (n + 127) // 128
    act = np.zeros((27, nt * 128), bool)
    act[:, :n] = got >= 0
    want = (act.reshape(27, nt, 128).any(2).astype(np.int64) << np.arange(27)[:, None]).sum(0)
    tm = srt.tile_mask[:nt].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    np.testing.assert_array_equal(tm, want)


@pytest.mark.parametrize("n,dims,B,stride,pad,per_group", [(6000, (9, 24, 24), 2, (2, 2, 2), (1, 1, 1), 0),
                                                           (4000, (8, 20, 20), 3, (2, 2, 2), (0, 1, 1), 2),
                                                           (50, (5, 6, 6), 1, (1, 2, 2), (1, 1, 1), 0)])
def test_rulebook_down_sorted_direct(n, dims, B, stride, pad, per_group):
    """Strided conv: the tile-sorted table built straight from the emitted output coordinates equals the oracle's
    table gathered through slot_row; output coordinates / count as in the natural-order call. Exact."""
    from uni3detr_b200 import ops
    coors = rand_coors(n, dims, B, n + 11)
    c = T(coors).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, n, B, dims)
    oc, on, ovm, srt, ocap = ops.rulebook_down(c, n_rows, n, vm, stride, pad, sorted_group=per_group)
    oc0, on0, _, nat_t, _ = ops.rulebook_down(c, n_rows, n, vm, stride, pad)
    m = int(on)
    assert m == int(on0) and torch.equal(oc[:m], oc0[:m])
    ref_coors, ref_nbr, _ = G.down_rulebook(coors, dims, stride, pad)
    np.testing.assert_array_equal(oc[:m].cpu().numpy(), ref_coors)
    slot_row = srt.slot_row[:m].cpu().numpy()
    np.testing.assert_array_equal(np.sort(slot_row), np.arange(m))
    np.testing.assert_array_equal(srt[:, :m].cpu().numpy(), ref_nbr[:, slot_row])
    np.testing.assert_array_equal(nat_t[:, :m].cpu().numpy(), ref_nbr)
    mask = ((ref_nbr >= 0).astype(np.int64) << np.arange(27)[:, None]).sum(0)
    grp = ref_coors[:, 0] // per_group if per_group else np.zeros(m, np.int64)
    key = grp[slot_row].astype(np.int64) * 4096 + _tile_key(mask)[slot_row]
    assert bool((np.diff(key) >= 0).all())
    nt = (m + 127) // 128
    act = np.zeros((27, nt * 128), bool)
    act[:, :m] = srt[:, :m].cpu().numpy() >= 0
    want = (act.reshape(27, nt, 128).any(2).astype(np.int64) << np.arange(27)[:, None]).sum(0)
    np.testing.assert_array_equal(srt.tile_mask[:nt].cpu().numpy().astype(np.int64) & 0xFFFFFFFF, want)


@pytest.mark.parametrize("seq_len,n_seq", [(300, 8), (900, 2), (5, 4), (129, 1)])
def test_mha_core_v_mn_major(seq_len, n_seq, monkeypatch):
    """U3D_MHA_VMN=1 stages V untransposed and uses an MN-major B operand for O = P V."""
    from uni3detr_b200 import ops
    monkeypatch.setenv("U3D_MHA_VMN", "1")
    g = torch.Generator().manual_seed(seq_len)
    heads, E = 8, 256
    qk = torch.randn(n_seq * seq_len, 2 * E, generator=g).bfloat16()
    v = torch.randn(n_seq * seq_len, E, generator=g).bfloat16()
    out = ops.mha_core(qk.to(DEV)[:, :E], qk.to(DEV)[:, E:], v.to(DEV), n_seq, seq_len, heads)

    def split(t):
        return t.float().view(n_seq, seq_len, heads, 32).permute(0, 2, 1, 3)
    ref = F.scaled_dot_product_attention(split(qk[:, :E]), split(qk[:, E:]), split(v))
    ref = ref.permute(0, 2, 1, 3).reshape(n_seq * seq_len, E)
    assert relerr(out, ref) < 1e-2


# ------------------------------------------------------------------ decoder GEMMs (tcgen05) ----
def _lin_ref(a, W, b, relu=False, mul=None, res1=None, res2=None, ln=None, relu_out=False):
    """fp32 torch reference of u3d_linear_tc on the same bf16 inputs."""
    y = a.float() @ W.to(torch.bfloat16).float().t()
    if b is not None:
        y = y + b.float()
    if relu:
        y = torch.relu(y)
    if mul is not None:
        y = y * mul.float()
    for r in (res1, res2):
        if r is not None:
            y = y + r.float()
    if ln is not None:
        y = torch.nn.functional.layer_norm(y, (y.shape[1],), ln[0].float(), ln[1].float(), ln[2])
    if relu_out:
        y = torch.relu(y)
    return y


@pytest.mark.parametrize("rows,K,N", [(300, 256, 256), (1000, 384, 256), (128 * 149 + 77, 256, 256),
                                       (513, 512, 256), (700, 256, 512), (129, 64, 64), (2000, 128, 128)])
def test_linear_tc_plain_and_relu(rows, K, N):
    """out = relu?(A W^T + b): K in {64..512}, N in {64, 128, 256, 512 (two passes)}, ragged last tile,
    more tiles than SMs (persistent loop + both TMEM accumulators + ring wrap-around)."""
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(rows + K + N)
    a = (torch.randn(rows, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    lin = ops.PackedLinear(W, b)
    for relu in (False, True):
        out = ops.linear_tc(a, lin, relu=relu)
        ref = _lin_ref(a, W, b, relu=relu)
        assert out.dtype == torch.bfloat16 and tuple(out.shape) == (rows, N)
        torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)
    # strided A view (column slice of a wider matrix, as the packed QK projection is consumed)
    wide = torch.zeros(rows, K + 64, dtype=torch.bfloat16, device=DEV)
    wide[:, 64:] = a
    out = ops.linear_tc(wide[:, 64:], lin)
    torch.testing.assert_close(out.float(), _lin_ref(a, W, b), rtol=1e-2, atol=1e-2)


def test_linear_tc_fused_epilogues():
    """Every epilogue stage: multiplier, two residuals, LayerNorm (+ReLU), second output, no bias."""
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(11)
    rows, K, N = 128 * 3 + 50, 256, 256
    bf = lambda *s: (torch.randn(*s, generator=g) * 0.7).to(torch.bfloat16).to(DEV)
    a, mul, r1, r2, add2 = bf(rows, K), bf(rows, N), bf(rows, N), bf(rows, N), bf(rows, N)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    gamma, beta = (1 + 0.2 * torch.randn(N, generator=g)).to(DEV), (0.3 * torch.randn(N, generator=g)).to(DEV)
    lin = ops.PackedLinear(W, b, ln=(gamma, beta, 1e-5))
    out = ops.linear_tc(a, lin, res1=r1, ln=True)                                  # out-proj + identity + LN
    torch.testing.assert_close(out.float(), _lin_ref(a, W, b, res1=r1, ln=(gamma, beta, 1e-5)), rtol=1e-2, atol=2e-2)
    out = ops.linear_tc(a, lin, res1=r1, res2=r2, ln=True)                         # cross-attn block
    torch.testing.assert_close(out.float(), _lin_ref(a, W, b, res1=r1, res2=r2, ln=(gamma, beta, 1e-5)),
                               rtol=1e-2, atol=2e-2)
    out = ops.linear_tc(a, lin, ln=True, relu_out=True)                            # Linear-LN-ReLU (cls branch)
    torch.testing.assert_close(out.float(), _lin_ref(a, W, b, ln=(gamma, beta, 1e-5), relu_out=True),
                               rtol=1e-2, atol=2e-2)
    out, out2 = ops.linear_tc(a, lin, mul=mul, add2=add2)                          # query_scale * qpos, x + qpos
    ref = _lin_ref(a, W, b, mul=mul)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(out2.float(), out.float() + add2.float(), rtol=1e-2, atol=2e-2)
    lin0 = ops.PackedLinear(W, None)
    torch.testing.assert_close(ops.linear_tc(a, lin0, relu=True).float(), _lin_ref(a, W, None, relu=True),
                               rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("N", [1, 8, 10])
def test_linear_tc_narrow_fp32_heads(N):
    """Final cls (10) / reg (8) / iou (1) layers: fp32 output, N padded to 16 inside; the reg variant
    also emits the refined reference points ref + (v0, v1, v4)."""
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(20 + N)
    rows, K = 1000, 256
    a = (torch.randn(rows, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    lin = ops.PackedLinear(W, b)
    ref = _lin_ref(a, W, b)
    if N == 8:
        r_in = torch.randn(rows, 3, generator=g).to(DEV)
        out, r_out = ops.linear_tc(a, lin, out_f32=True, ref_in=r_in)
        torch.testing.assert_close(r_out, r_in + torch.stack((out[:, 0], out[:, 1], out[:, 4]), 1), rtol=0, atol=1e-6)
    else:
        out = ops.linear_tc(a, lin, out_f32=True)
    assert out.dtype == torch.float32 and tuple(out.shape) == (rows, N)
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3)


def test_pos3_ln_relu():
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(3)
    ref = torch.randn(777, 3, generator=g).to(DEV)
    W, b = torch.randn(256, 3, generator=g).to(DEV), torch.randn(256, generator=g).to(DEV)
    gamma, beta = (1 + 0.1 * torch.randn(256, generator=g)).to(DEV), (0.1 * torch.randn(256, generator=g)).to(DEV)
    want = torch.relu(torch.nn.functional.layer_norm(ref @ W.t() + b, (256,), gamma, beta, 1e-5))
    torch.testing.assert_close(ops.pos3_ln_relu(ref, W, b, gamma, beta, 1e-5, torch.float32), want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ops.pos3_ln_relu(ref, W, b, gamma, beta, 1e-5, torch.bfloat16).float(), want,
                               rtol=1e-2, atol=1e-2)


# ------------------------------------------------------- fp32 sparse conv on tensor cores (3xBF16) ----
@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 32), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256)])
def test_spconv_x3_fp32_accuracy(cin, cout):
    """u3d_spconv_fwd_packed_x3 (fp32 layers as x_hi*w_hi + x_hi*w_lo + x_lo*w_hi on tcgen05) against the fp32
    oracle: 1e-4 max-norm relative - an order of magnitude inside the 1e-3 fp32 bound of north_star, where the plain
    bf16 kernel sits at ~4e-3. Covers Cout > 128 (several 128-channel launches), residual + ReLU, the hi/lo output
    pair feeding a second conv, and a tile-sorted rulebook."""
    from uni3detr_b200 import ops
    dims, B, n = (8, 24, 24), 2, 3000
    coors, x, w, scale, shift = conv_case(n, dims, B, cin, cout, cin * 1000 + cout + 7)
    nbr_ref = G.subm_rulebook(coors, dims)
    c = T(coors).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    vm = ops.voxmap_build(c, n_rows, n, B, dims)
    nbr = ops.rulebook_subm(c, n_rows, n, vm)
    pk = ops.PackedConvX3(w.reshape(27, cin, cout).to(DEV))
    res = torch.randn(n, cout, generator=torch.Generator().manual_seed(1))
    x2 = ops.split_bf16(x.to(DEV))
    assert relerr(ops.merge_bf16(x2), x) < 2e-5                       # the [hi | lo] pair carries ~16 bits
    for residual, relu in [(None, True), (res, True), (None, False)]:
        ref = oracle_conv(x, nbr_ref, w, n, scale, shift, residual, relu)
        r2 = None if residual is None else ops.split_bf16(residual.to(DEV))
        y2 = ops.spconv_fwd_packed_x3(x2, nbr, n_rows, n, pk, scale.to(DEV), shift.to(DEV), residual=r2, relu=relu)
        assert tuple(y2.shape) == (n, 2 * cout)
        assert relerr(ops.merge_bf16(y2), ref) < 1e-4, (residual is not None, relu, relerr(ops.merge_bf16(y2), ref))
    if cin <= 32:                                                     # tile-sorted table (slot_row path)
        srt = ops.rulebook_sort_tiles(nbr, n_rows, n)
        y2 = ops.spconv_fwd_packed_x3(x2, srt, n_rows, n, pk, scale.to(DEV), shift.to(DEV), relu=True)
        assert relerr(ops.merge_bf16(y2), oracle_conv(x, nbr_ref, w, n, scale, shift, None, True)) < 1e-4


def test_spconv_x3_pointwise_identity_table():
    """1x1x1 conv (conv_out) through the gather kernel with an identity rulebook."""
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, cin, cout = 2777, 128, 256
    x = torch.randn(n, cin, generator=g)
    w = torch.randn(1, cin, cout, generator=g) / cin ** 0.5
    scale, shift = 1 + 0.1 * torch.randn(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    pk = ops.PackedConvX3(w.to(DEV))
    y2 = ops.spconv_fwd_packed_x3(ops.split_bf16(x.to(DEV)), ops.identity_rulebook(n, DEV), n_rows, n, pk, scale.to(DEV),
                                  shift.to(DEV), relu=True)
    ref = torch.relu((x @ w[0]) * scale + shift)
    assert relerr(ops.merge_bf16(y2), ref) < 1e-4


@pytest.mark.parametrize("cin,cout", [(64, 64), (32, 32), (32, 64), (16, 32), (64, 128), (128, 128), (16, 16)])
def test_spconv_switches_reverse_tiles_and_deep_ring(monkeypatch, cin, cout):
    """Instruction-form / ring variants of the rows-on-N kernel give the same result as the default launch (the
    weight-stationary `tcgen05.mma.ws` form with M = Cout for Cout = 32 / 64): reverse tile order (`reverse=True`), the
    plain M = 64 form (U3D_TN_WS=0), the replicated M = 128 form (+ U3D_TN_M64=0), the single-buffered-slice 4-stage
    ring (+ U3D_TN_SLICE_BUFS=1), per-stage rulebook rows on / off (U3D_TN_RING), bf16 and the 3xBF16 fp32 form; and
    the default agrees with the fp32 reference."""
    from uni3detr_b200 import ops
    dims, B, n = (8, 24, 24), 2, 3000
    coors, x, w, scale, shift = conv_case(n, dims, B, cin, cout, 4242)
    c = T(coors).to(DEV)
    n_rows = torch.tensor([n], dtype=torch.int32, device=DEV)
    nbr = ops.rulebook_subm(c, n_rows, n, ops.voxmap_build(c, n_rows, n, B, dims))
    wd = w.reshape(27, cin, cout).to(DEV)
    wp = ops.spconv_pack_weights(wd.bfloat16().contiguous())
    pk = ops.PackedConvX3(wd)
    xb, x2 = x.to(DEV).bfloat16(), ops.split_bf16(x.to(DEV))

    def run(**kw):
        a = ops.spconv_fwd_packed(xb, nbr, n_rows, n, wp, 27, cin, cout, scale.to(DEV), shift.to(DEV), relu=True, **kw)
        b = ops.spconv_fwd_packed_x3(x2, nbr, n_rows, n, pk, scale.to(DEV), shift.to(DEV), relu=True, **kw)
        return a.clone(), b.clone()
    base = run()
    ref = oracle_conv(x, G.subm_rulebook(coors, dims), w, n, scale, shift, None, True)
    assert relerr(base[0][:n], ref) < 1e-2
    assert relerr(ops.merge_bf16(base[1])[:n], ref) < 1e-4
    for name, env, kw in (("reverse", {}, dict(reverse=True)), ("m64", {"U3D_TN_WS": "0"}, {}),
                          ("stage_rows", {"U3D_TN_RING": "1"}, {}), ("tile_slices", {"U3D_TN_RING": "0"}, {}),
                          ("stage_rows_m64", {"U3D_TN_RING": "1", "U3D_TN_WS": "0"}, {}),
                          ("m128", {"U3D_TN_WS": "0", "U3D_TN_M64": "0"}, {}),
                          ("deep", {"U3D_TN_WS": "0", "U3D_TN_SLICE_BUFS": "1", "U3D_TN_M64": "0"}, {})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got = run(**kw)
        for k in env:
            monkeypatch.delenv(k)
        for g_, b_ in zip(got, base):
            assert torch.equal(g_, b_), name     # same accumulation order per row: bit-identical
