"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product path (uni3detr_b200/).

CPU restatement (numpy) of the integer/geometry half of the Uni3DETR forward hot path:
voxelization, simple VFEs, sparse-conv rulebooks and D-FPS.

PARITY STATUS: **unpinned by the reference** for everything in this file. The arithmetic lives
in third-party packages that are not vendored under /root/reference and cannot be installed
here (mmcv-full 1.x `_ext`: hard/dynamic voxelize, DynamicScatter, furthest_point_sample,
gather_points; spconv: get_indice_pairs) and the reference ships no tests or golden vectors
(SURVEY.md §4, §8c). Each function restates the published algorithm as recorded in
SURVEY.md Appendix A and cites the reference call site it serves. Independent known-answer
checks (np.unique, dense conv3d equivalence, brute force) live in tests/test_oracle.py.
"""
import numpy as np

F32 = np.float32


def grid_size_xyz(pc_range, voxel_size):
    """mmcv Voxelization.__init__: round((range[3:]-range[:3])/voxel_size) in fp32 -> (x,y,z)."""
    r = np.asarray(pc_range, F32)
    v = np.asarray(voxel_size, F32)
    return np.round((r[3:] - r[:3]) / v).astype(np.int64)


def point_cells(points, pc_range, voxel_size):
    """c = floor((p - lo)/vs) in fp32 per axis; valid iff 0 <= c < grid."""
    pts = np.asarray(points, F32)
    lo = np.asarray(pc_range[:3], F32)
    vs = np.asarray(voxel_size, F32)
    grid = grid_size_xyz(pc_range, voxel_size)
    c = np.floor((pts[:, :3] - lo) / vs)
    valid = np.all((c >= 0) & (c < grid.astype(F32)), axis=1)
    ci = np.where(valid[:, None], c, -1).astype(np.int64)  # x,y,z
    return ci, valid, grid


def hard_voxelize(points, pc_range, voxel_size, max_pts, max_voxels, deterministic=True):
    """mmcv.ops.Voxelization (hard) for ONE sample — call site detectors/uni3detr.py:148 via
    MVXTwoStageDetector.voxelize. Returns voxels (M,max_pts,C) zero padded in arrival order,
    coors (M,3) zyx int32, num_points (M,) int32. deterministic=True: voxels in
    first-appearance order; False: ascending linear index (canonicalised)."""
    pts = np.asarray(points, F32)
    C = pts.shape[1]
    ci, valid, grid = point_cells(pts, pc_range, voxel_size)
    idx = np.nonzero(valid)[0]
    if idx.size == 0:
        return (np.zeros((0, max_pts, C), F32), np.zeros((0, 3), np.int32), np.zeros((0,), np.int32))
    c = ci[idx]
    lin = (c[:, 2] * grid[1] + c[:, 1]) * grid[0] + c[:, 0]
    uniq, first, inv = np.unique(lin, return_index=True, return_inverse=True)
    if deterministic:
        order = np.argsort(first, kind="stable")
    else:
        order = np.arange(len(uniq))
    vid_of_uniq = np.empty(len(uniq), np.int64)
    vid_of_uniq[order] = np.arange(len(uniq))
    vid = vid_of_uniq[inv]
    M = len(uniq) if max_voxels <= 0 else min(len(uniq), int(max_voxels))
    srt = np.argsort(vid, kind="stable")           # arrival order inside each voxel
    vs_sorted = vid[srt]
    starts = np.searchsorted(vs_sorted, np.arange(len(uniq)))
    rank = np.arange(len(srt)) - starts[vs_sorted]
    keep = (vs_sorted < M) & (rank < max_pts)
    voxels = np.zeros((M, max_pts, C), F32)
    voxels[vs_sorted[keep], rank[keep]] = pts[idx[srt[keep]]]
    num = np.bincount(vs_sorted[keep], minlength=M).astype(np.int32)
    coors = np.zeros((M, 3), np.int32)
    first_pt = c[first[order[:M]]]
    coors[:, 0], coors[:, 1], coors[:, 2] = first_pt[:, 2], first_pt[:, 1], first_pt[:, 0]
    return voxels, coors, num


def hard_simple_vfe(voxels, num_points, num_features):
    """mmdet3d HardSimpleVFE (call site uni3detr.py:149): sum over the padded axis / count."""
    s = np.zeros((voxels.shape[0], num_features), F32)
    for j in range(voxels.shape[1]):              # sequential fp32 accumulation
        s = s + voxels[:, j, :num_features]
    return (s / num_points.astype(F32)[:, None]).astype(F32)


def voxelize_batch_hard(points_list, pc_range, voxel_size, max_pts, max_voxels, deterministic=True,
                        num_features=None):
    """MVXTwoStageDetector.voxelize: per-sample voxelize, left-pad coors with the batch index,
    concatenate (uni3detr.py:148). Returns voxels, num_points, coors (M,4), feats (VFE mean)."""
    vs, ns, cs = [], [], []
    for b, p in enumerate(points_list):
        v, c, n = hard_voxelize(p, pc_range, voxel_size, max_pts, max_voxels, deterministic)
        vs.append(v)
        ns.append(n)
        cs.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
    voxels, num, coors = np.concatenate(vs), np.concatenate(ns), np.concatenate(cs)
    nf = num_features or voxels.shape[2]
    return voxels, num, coors, hard_simple_vfe(voxels, num, nf)


def dynamic_voxelize(points, pc_range, voxel_size):
    """mmcv.ops.Voxelization(max_num_points=-1): per-point (N,3) zyx, -1 rows when out of range
    (call site uni3detr.py:157-159)."""
    ci, valid, _ = point_cells(points, pc_range, voxel_size)
    return ci[:, ::-1].astype(np.int32)


def dynamic_scatter_mean(features, coors):
    """mmdet3d DynamicSimpleVFE -> mmcv DynamicScatter(mean) (call site uni3detr.py:167):
    unique over (b,z,y,x) sorted lexicographically, rows with a negative coord dropped,
    per-voxel mean of ALL its points. Returns feats (M,C) fp32, coors (M,4) int32."""
    feats = np.asarray(features, F32)
    coors = np.asarray(coors, np.int64)
    ok = np.all(coors >= 0, axis=1)
    uniq, inv, cnt = np.unique(coors[ok], axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    out = np.zeros((len(uniq), feats.shape[1]), np.float64)
    np.add.at(out, inv, feats[ok].astype(np.float64))
    return (out / cnt[:, None]).astype(F32), uniq.astype(np.int32)


# ---------------------------------------------------------------- rulebook ----
def _lin(coors, dims):
    D, H, W = dims
    c = coors.astype(np.int64)
    return ((c[:, 0] * D + c[:, 1]) * H + c[:, 2]) * W + c[:, 3]


class _Lookup:
    def __init__(self, coors, dims):
        self.dims = dims
        lin = _lin(coors, dims)
        self.order = np.argsort(lin, kind="stable")
        self.sorted = lin[self.order]

    def rows(self, coords, valid):
        """row index of each (b,z,y,x) in coords (or -1)."""
        out = np.full(len(coords), -1, np.int64)
        if len(self.sorted) == 0:
            return out
        q = _lin(np.where(valid[:, None], coords, 0), self.dims)
        pos = np.searchsorted(self.sorted, q)
        pos = np.minimum(pos, len(self.sorted) - 1)
        hit = valid & (self.sorted[pos] == q)
        out[hit] = self.order[pos[hit]]
        return out


def conv_out_dims(in_dims, stride, pad, k=3):
    return tuple((int(d) + 2 * int(p) - k) // int(s) + 1 for d, s, p in zip(in_dims, stride, pad))


def neighbour_table(out_coors, in_coors, in_dims, stride=(1, 1, 1), pad=(1, 1, 1)):
    """Output-stationary rulebook: nbr[k, o] = input row at out*stride - pad + k_off (or -1),
    k = (kz*3+ky)*3+kx. Equivalent to spconv's indice_pairs: the pair list of
    offset k is {(nbr[k,o], o) : nbr[k,o] >= 0}."""
    lk = _Lookup(in_coors, in_dims)
    nbr = np.full((27, len(out_coors)), -1, np.int64)
    oc = out_coors.astype(np.int64)
    for k in range(27):
        off = np.array([k // 9, (k // 3) % 3, k % 3])
        c = oc.copy()
        c[:, 1:] = oc[:, 1:] * np.asarray(stride) - np.asarray(pad) + off
        valid = np.all((c[:, 1:] >= 0) & (c[:, 1:] < np.asarray(in_dims)), axis=1)
        nbr[k] = lk.rows(c, valid)
    return nbr.astype(np.int32)


def subm_rulebook(coors, dims):
    """SubMConv3d(k=3): outputs == inputs (same order); instantiated at
    sparse_encoder_hd.py:71-88,193-199."""
    return neighbour_table(coors, coors, dims)


def down_rulebook(in_coors, in_dims, stride, pad):
    """SparseConv3d(k=3,stride,pad) (sparse_encoder_hd.py:181-192): active outputs =
    unique{(i + p - k)/s : divisible, in range}, ordered by ascending linear index."""
    out_dims = conv_out_dims(in_dims, stride, pad)
    ic = in_coors.astype(np.int64)
    cands = []
    for k in range(27):
        off = np.array([k // 9, (k // 3) % 3, k % 3])
        t = ic[:, 1:] + np.asarray(pad) - off
        ok = np.all((t >= 0) & (t % np.asarray(stride) == 0), axis=1)
        o = t // np.asarray(stride)
        ok &= np.all(o < np.asarray(out_dims), axis=1)
        cands.append(np.concatenate([ic[ok, :1], o[ok]], 1))
    cand = np.concatenate(cands) if cands else np.zeros((0, 4), np.int64)
    if len(cand):
        lin = _lin(cand, out_dims)
        _, first = np.unique(lin, return_index=True)      # sorted ascending by linear index
        out_coors = cand[first]
    else:
        out_coors = np.zeros((0, 4), np.int64)
    nbr = neighbour_table(out_coors, in_coors, in_dims, stride, pad)
    return out_coors.astype(np.int32), nbr, out_dims


def pairs_from_table(nbr):
    """spconv-1.x view: list over k of (in_rows, out_rows)."""
    return [(nbr[k][nbr[k] >= 0].astype(np.int64), np.nonzero(nbr[k] >= 0)[0]) for k in range(len(nbr))]


# --------------------------------------------------------------------- FPS ----
def fps_tie_key(n, tie_block=1024):
    """Order in which mmcv's furthest_point_sample kernel resolves exact distance ties (smaller wins). The kernel
    runs bs = min(tie_block, 2^floor(log2 n)) threads; thread t scans k = t, t+bs, ... keeping its first maximum
    (strict >), then a shared-memory tree (s = bs/2 .. 1: slot t takes slot t+s only if strictly greater) picks the
    block winner: two tied candidates meet at the level of the lowest bit in which their thread ids differ and the
    one with that bit clear survives, i.e. the smallest BIT-REVERSED thread id wins, then the lowest k.
    (mmcv/ops/csrc/common/cuda/furthest_point_sample_cuda_kernel.cuh - recalled, the source is not vendored;
    literal restatement: fps_block_reference below.) The kernel's launcher takes the exponent as
    int(log(n) / log(2.0)) in double precision, which equals floor(log2 n) for every n up to 2^20 including the exact
    powers of two (tests/test_oracle.py). tie_block=0: plain lowest index."""
    k = np.arange(n, dtype=np.int64)
    if not tie_block or n == 0:
        return k
    lg = min(int(np.floor(np.log2(n))), int(np.log2(tie_block)))
    bs = 1 << lg
    r = k % bs
    rev = np.zeros(n, np.int64)
    for b in range(lg):
        rev |= ((r >> b) & 1) << (lg - 1 - b)
    return (rev << 12) | (k // bs)


def fps_block_reference(temp, bs):
    """Literal restatement of the arg-max of one FPS iteration as the reference kernel computes it (per-thread strided
    scan + shared-memory tree), for pinning fps_tie_key: returns the selected index."""
    n = len(temp)
    dists = np.full(bs, -1.0, F32)
    dists_i = np.zeros(bs, np.int64)
    for t in range(bs):
        best, besti = F32(-1.0), 0
        for k in range(t, n, bs):
            if temp[k] > best:
                best, besti = temp[k], k
        dists[t], dists_i[t] = best, besti
    s = bs // 2
    while s >= 1:
        for t in range(s):
            if dists[t + s] > dists[t]:
                dists[t], dists_i[t] = dists[t + s], dists_i[t + s]
        s //= 2
    return int(dists_i[0])


def furthest_point_sample(xyz, npoint, tie_block=1024):
    """mmcv furthest_point_sample / D-FPS (call sites uni3detr.py:138,179,184):
    idx[0]=0, temp=1e10, d = ((dx*dx+dy*dy)+dz*dz) in fp32 without FMA, exact ties by fps_tie_key."""
    p = np.ascontiguousarray(xyz, F32)
    n = len(p)
    idx = np.zeros(npoint, np.int32)
    if n == 0:
        return idx
    temp = np.full(n, 1e10, F32)
    key = fps_tie_key(n, tie_block)
    last = 0
    for j in range(1, npoint):
        d = p - p[last]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        temp = np.minimum(temp, d2)
        cand = np.nonzero(temp == temp.max())[0]
        last = int(cand[np.argmin(key[cand])]) if len(cand) > 1 else int(cand[0])
        idx[j] = last
    return idx


def fps_input_view(points, stride_quirk=True):
    """the reference hands the (1,N,C) tensor to a kernel that strides by 3 floats."""
    pts = np.ascontiguousarray(points, F32)
    n = len(pts)
    if stride_quirk and pts.shape[1] != 3:
        return pts.reshape(-1)[:3 * n].reshape(n, 3)
    return pts[:, :3]


def shift_scale_unit(sampled):
    """shift_scale_points with dst range [0,1] (uni3detr.py:18-46,181,187):
    ((p - min) * 1) / (max - min) + 0 over the sampled set of one scene."""
    s = np.asarray(sampled, F32)
    mn, mx = s.min(0), s.max(0)
    return ((s - mn) * F32(1.0)) / (mx - mn) + F32(0.0)


def fps_queries(points_list, coors_per_scene, nq, stride_quirk=True):
    """uni3detr.py:178-189: FPS on raw points and on voxel coordinates -> (B, 2nq, 3) in [0,1].
    coors_per_scene[b]: (M_b,3) zyx coordinates (float or int)."""
    out = []
    for pts, cz in zip(points_list, coors_per_scene):
        pts = np.asarray(pts, F32)
        i1 = furthest_point_sample(fps_input_view(pts, stride_quirk), nq)
        a = shift_scale_unit(pts[i1, :3])
        czf = np.asarray(cz, F32)
        i2 = furthest_point_sample(czf, nq)
        b = shift_scale_unit(czf[i2][:, [2, 1, 0]])
        out.append(np.concatenate([a, b], 0))
    return np.stack(out).astype(F32)
