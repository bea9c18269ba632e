"""Independent known-answer checks of the oracle's restatement of the THIRD-PARTY ops
(mmcv voxelize / DynamicScatter / FPS, spconv rulebook + indice_conv, mmcv transformer bricks):
their sources are not vendored under the reference, so parity there is unpinned; these tests
cross-check the restatement against brute-force loops and torch library ops."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import geometry as G
from oracle import model as M

PCR = [-3.2, -0.2, -2.0, 3.2, 6.2, 0.56]
VS = [0.02, 0.02, 0.02]


def cloud(n, seed, C=4, spread=1.0):
    rng = np.random.default_rng(seed)
    lo, hi = np.array(PCR[:3]), np.array(PCR[3:])
    p = lo + rng.random((n, 3)) * (hi - lo) * spread
    p[: n // 10] += 10.0  # some points out of range
    out = np.zeros((n, C), np.float32)
    out[:, :3] = p
    out[:, 3:] = rng.random((n, C - 3))
    return out


def brute_hard_voxelize(points, pcr, vs, max_pts, max_voxels):
    """mmcv hard_voxelize CPU semantics as a plain python loop."""
    grid = G.grid_size_xyz(pcr, vs)
    lo, v = np.float32(pcr[:3]), np.float32(vs)
    vox_of = {}
    voxels, coors, num = [], [], []
    for p in points:
        c = np.floor((p[:3] - lo) / v)
        if np.any(c < 0) or np.any(c >= grid):
            continue
        key = (int(c[2]), int(c[1]), int(c[0]))
        if key not in vox_of:
            if max_voxels > 0 and len(voxels) >= max_voxels:
                continue
            vox_of[key] = len(voxels)
            voxels.append(np.zeros((max_pts, points.shape[1]), np.float32))
            coors.append(key)
            num.append(0)
        i = vox_of[key]
        if num[i] < max_pts:
            voxels[i][num[i]] = p
            num[i] += 1
    return np.array(voxels), np.array(coors, np.int32).reshape(-1, 3), np.array(num, np.int32)


def test_hard_voxelize_vs_bruteforce():
    for seed, n, mp, mv, spread in [(0, 3000, 5, 40000, 0.1), (1, 3000, 2, 100, 0.05), (2, 500, 5, 0, 1.0)]:
        pts = cloud(n, seed, spread=spread)
        v, c, k = G.hard_voxelize(pts, PCR, VS, mp, mv)
        bv, bc, bk = brute_hard_voxelize(pts, PCR, VS, mp, mv)
        np.testing.assert_array_equal(c, bc)
        np.testing.assert_array_equal(k, bk)
        np.testing.assert_array_equal(v, bv)


def test_hard_voxelize_empty_and_all_out_of_range():
    v, c, k = G.hard_voxelize(np.zeros((0, 4), np.float32), PCR, VS, 5, 10)
    assert v.shape == (0, 5, 4) and c.shape == (0, 3) and k.shape == (0,)
    pts = np.full((7, 4), 100.0, np.float32)
    v, c, k = G.hard_voxelize(pts, PCR, VS, 5, 10)
    assert len(v) == 0


def test_dynamic_scatter_vs_torch_unique():
    pts = cloud(4000, 3, spread=0.08)
    c = G.dynamic_voxelize(pts, PCR, VS)
    cb = np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)
    feats, coors = G.dynamic_scatter_mean(pts, cb)
    ok = (c >= 0).all(1)
    u, inv = torch.unique(torch.from_numpy(cb[ok].astype(np.int64)), dim=0, return_inverse=True)
    np.testing.assert_array_equal(coors, u.numpy())
    ref = torch.zeros(len(u), 4, dtype=torch.float64).index_add_(0, inv, torch.from_numpy(pts[ok]).double())
    cnt = torch.bincount(inv, minlength=len(u)).double()
    np.testing.assert_allclose(feats, (ref / cnt[:, None]).numpy(), rtol=1e-6, atol=1e-6)
    assert (c[~ok] == -1).all()


def rand_coors(n, dims, B, seed):
    rng = np.random.default_rng(seed)
    D, H, W = dims
    lin = rng.choice(B * D * H * W, size=n, replace=False)
    c = np.stack([lin // (D * H * W), (lin // (H * W)) % D, (lin // W) % H, lin % W], 1).astype(np.int32)
    return c


def densify(x, coors, B, dims):
    d = torch.zeros((B, x.shape[1]) + tuple(dims))
    c = torch.from_numpy(coors.astype(np.int64))
    d[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = x
    return d


def test_subm_conv_equals_dense_conv3d_masked():
    dims, B = (6, 9, 8), 2
    coors = rand_coors(300, dims, B, 5)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(300, 5, generator=g)
    w = torch.randn(3, 3, 3, 5, 7, generator=g)
    nbr = G.subm_rulebook(coors, dims)
    y = M.sparse_conv(x, nbr, w, 300)
    dense = F.conv3d(densify(x, coors, B, dims), w.permute(4, 3, 0, 1, 2), padding=1)
    c = torch.from_numpy(coors.astype(np.int64))
    ref = dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]]
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)


def test_down_conv_equals_dense_conv3d():
    for dims, stride, pad in [((7, 10, 9), (2, 2, 2), (1, 1, 1)), ((8, 9, 9), (2, 2, 2), (0, 1, 1)),
                              ((6, 8, 8), (2, 2, 2), (0, 0, 0))]:
        B = 2
        coors = rand_coors(200, dims, B, 7)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(200, 4, generator=g)
        w = torch.randn(3, 3, 3, 4, 6, generator=g)
        oc, nbr, od = G.down_rulebook(coors, dims, stride, pad)
        y = M.sparse_conv(x, nbr, w, len(oc))
        dense = F.conv3d(densify(x, coors, B, dims), w.permute(4, 3, 0, 1, 2), stride=stride, padding=pad)
        assert tuple(dense.shape[2:]) == tuple(od)
        # active outputs == sites whose receptive field holds an active input
        occ = F.conv3d(densify(torch.ones(200, 1), coors, B, dims), torch.ones(1, 1, 3, 3, 3),
                       stride=stride, padding=pad)[:, 0] > 0
        act = torch.nonzero(occ).numpy()
        np.testing.assert_array_equal(oc, act.astype(np.int32))       # ascending linear order
        c = torch.from_numpy(oc.astype(np.int64))
        torch.testing.assert_close(y, dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]], rtol=1e-4, atol=1e-4)


def test_rulebook_bruteforce():
    dims, B = (5, 6, 7), 2
    coors = rand_coors(120, dims, B, 9)
    nbr = G.subm_rulebook(coors, dims)
    table = {tuple(c): i for i, c in enumerate(coors.tolist())}
    for k in range(27):
        off = (k // 9 - 1, (k // 3) % 3 - 1, k % 3 - 1)
        for o, c in enumerate(coors.tolist()):
            q = (c[0], c[1] + off[0], c[2] + off[1], c[3] + off[2])
            assert nbr[k, o] == table.get(q, -1)
    assert (nbr[13] == np.arange(120)).all()     # centre offset maps every site to itself


def test_fps_bruteforce_and_ties():
    rng = np.random.default_rng(4)
    p = rng.random((257, 3)).astype(np.float32)
    idx = G.furthest_point_sample(p, 40)
    # brute force in float64 on the same fp32 inputs must agree when no near-ties exist
    temp = np.full(257, 1e10)
    last, ref = 0, [0]
    for _ in range(39):
        d = ((p.astype(np.float64) - p[last]) ** 2).sum(1)
        temp = np.minimum(temp, d)
        last = int(np.argmax(temp))
        ref.append(last)
    assert idx.tolist() == ref
    # integer lattice: exact ties everywhere -> lowest index wins
    lat = np.array([[z, y, x] for z in range(3) for y in range(3) for x in range(3)], np.float32)
    idx = G.furthest_point_sample(lat, 5, tie_block=0)
    assert idx[0] == 0 and idx[1] == 26           # unique farthest corner
    assert idx[2] == 5                            # (0,1,2),(0,2,1),(1,0,2).. all at min-dist 5 -> lowest index
    assert len(set(idx.tolist())) == 5
    # reference tie order (block of 16 threads for 27 points): candidates 5,7,11,15,19,21 -> thread ids 5,7,11,15,3,5;
    # bit-reversed (4 bits) 10,14,13,15,12,10 -> thread 5 wins, its first maximum is k = 5
    idx = G.furthest_point_sample(lat, 5)
    assert idx[:3].tolist() == [0, 26, 5] and len(set(idx.tolist())) == 5


def test_fps_tie_order_matches_block_reduction():
    """The closed-form tie key (bit-reversed thread id, then index) against a literal restatement of the reference
    kernel's arg-max: per-thread strided scan with strict '>' + shared-memory tree that keeps the lower slot."""
    rng = np.random.default_rng(11)
    for n, cap in ((27, 1024), (100, 1024), (1000, 64), (2500, 1024), (5000, 512), (3, 1024), (1, 1024)):
        bs = min(cap, 1 << int(np.floor(np.log2(n))))
        key = G.fps_tie_key(n, cap)
        assert len(set(key.tolist())) == n                     # a total order
        for trial in range(6):
            temp = rng.integers(0, 4, n).astype(np.float32)    # few distinct values: ties everywhere
            if trial == 5:
                temp[:] = 0.0
            want = G.fps_block_reference(temp, bs)
            cand = np.nonzero(temp == temp.max())[0]
            assert int(cand[np.argmin(key[cand])]) == want, (n, cap, trial)
    np.testing.assert_array_equal(G.fps_tie_key(10, 0), np.arange(10))
    # the reference launcher's exponent int(log(n) / log(2.0)) (double) == floor(log2 n), exact powers of two included
    import math
    for n in list(range(1, 5000)) + [2 ** k + d for k in range(12, 21) for d in (-1, 0, 1)]:
        assert int(math.log(float(n)) / math.log(2.0)) == int(np.floor(np.log2(n))) == n.bit_length() - 1, n


def test_fps_stride_quirk_view():
    pts = np.arange(40, dtype=np.float32).reshape(10, 4)
    v = G.fps_input_view(pts, True)
    np.testing.assert_array_equal(v, np.arange(30, dtype=np.float32).reshape(10, 3))
    np.testing.assert_array_equal(G.fps_input_view(pts, False), pts[:, :3])
    np.testing.assert_array_equal(G.fps_input_view(pts[:, :3].copy(), True), pts[:, :3])


def test_transformer_layer_vs_torch_modules():
    """mmcv BaseTransformerLayer wiring against nn.MultiheadAttention / nn modules."""
    torch.manual_seed(0)
    E, nq, B = 256, 7, 2
    mha = torch.nn.MultiheadAttention(E, 8).eval()
    x, pos = torch.randn(nq, B, E), torch.randn(nq, B, E)
    with torch.no_grad():
        ref = x + mha(x + pos, x + pos, x)[0]
    sd = {"attentions.0.attn." + k: v for k, v in mha.state_dict().items()}
    qk = x + pos
    out, _ = F.multi_head_attention_forward(
        qk, qk, x, E, 8, sd["attentions.0.attn.in_proj_weight"], sd["attentions.0.attn.in_proj_bias"],
        None, None, False, 0.0, sd["attentions.0.attn.out_proj.weight"],
        sd["attentions.0.attn.out_proj.bias"], training=False, need_weights=False)
    torch.testing.assert_close(x + out, ref, rtol=1e-5, atol=1e-5)


def test_bev_iou_closed_forms():
    """oracle/postproc.py rotated IoU against analytic cases."""
    from oracle import postproc as PP
    a = [0, 0, 0, 2, 2, 1, 0.0]
    assert abs(PP.bev_iou(a, a) - 1.0) < 1e-12
    assert abs(PP.bev_iou(a, [1, 0, 0, 2, 2, 1, 0.0]) - (2.0 / 6.0)) < 1e-12        # half overlap
    assert PP.bev_iou(a, [5, 5, 0, 2, 2, 1, 0.3]) == 0.0
    # unit square vs itself rotated 45 deg: intersection = regular octagon, area 2*(sqrt2 - 1) * 1
    sq = [0, 0, 0, 1, 1, 1, 0.0]
    inter = 2 * (np.sqrt(2) - 1)
    assert abs(PP.bev_iou(sq, [0, 0, 0, 1, 1, 1, np.pi / 4]) - inter / (2 - inter)) < 1e-12
    # heading periodicity and dx/dy swap at 90 deg
    assert abs(PP.bev_iou([0, 0, 0, 4, 2, 1, np.pi / 2], [0, 0, 0, 2, 4, 1, 0.0]) - 1.0) < 1e-12
    # greedy order: the best box suppresses its neighbour, the far one survives
    boxes = np.array([[0, 0, 0, 2, 2, 1, 0], [0.2, 0, 0, 2, 2, 1, 0], [5, 0, 0, 2, 2, 1, 0]], float)
    np.testing.assert_array_equal(PP.nms3d(boxes, [0.5, 0.9, 0.1], 0.5), [1, 2])
