"""Tensor-level wrappers over the C ABI (include/u3d.h).

PyTorch is plumbing here: it owns device memory and the stream. Every function
enqueues on ``torch.cuda.current_stream()`` and never synchronises; data-dependent
sizes stay on the device (int32 counters), buffers are sized by capacity.
There is no CPU path: tensors must live on a CUDA device.
"""
import ctypes
import functools
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib

# -- optional per-op device timing (bench.py's roofline pass): CUDA events on the launching stream
_PROF = None


def profile_begin():
    global _PROF
    _PROF = []


def profile_end():
    """Returns [(op name, info dict, milliseconds)] for every op since profile_begin()."""
    global _PROF
    rec, _PROF = _PROF or [], None
    torch.cuda.synchronize()
    return [(n, i, s.elapsed_time(e)) for n, i, s, e in rec]


def _live_pairs(nbr, n_out):
    n = int(n_out)
    return n if nbr is None else int((nbr[:, :n] >= 0).sum())


def _timed(info=None):
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*a, **k):
            if _PROF is None:
                return fn(*a, **k)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.current_stream().synchronize()   # each op is timed alone (profile mode only)
            # keep the GPU busy (~150 us spin) while the host allocates / marshals / launches, so that the two
            # events bracket the op's kernels only and not the Python time between them
            torch.cuda._sleep(300000)
            s.record()
            r = fn(*a, **k)
            e.record()
            # info() returns plain python numbers (it may synchronise); no tensor is kept alive, so
            # the profile pass does not perturb the caching allocator
            _PROF.append((fn.__name__, info(r, *a, **k) if info else {}, s, e))
            return r
        return wrapper
    return deco

U3D_F32, U3D_BF16 = 0, 1
ORDER_FIRST_APPEARANCE, ORDER_LINEAR = 0, 1
_DT = {torch.float32: U3D_F32, torch.bfloat16: U3D_BF16}


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _farr(vals):
    arr = (ctypes.c_float * len(vals))(*[float(v) for v in vals])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def _iarr(vals):
    arr = (ctypes.c_int32 * len(vals))(*[int(v) for v in vals])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def _req(t: torch.Tensor, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise _lib.U3DError(f"{name} must be a CUDA tensor: libu3d_b200 has no CPU path")
    if not t.is_contiguous():
        raise _lib.U3DError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise _lib.U3DError(f"{name} must be {dtype}, got {t.dtype}")
    return t


@dataclass
class VoxelMap:
    """Coordinate index of one resolution level (see include/u3d.h)."""
    words: torch.Tensor            # (nwords, 2) int32 = uint2 {bits, prefix}
    perm: Optional[torch.Tensor]   # (ranks,) int32 rank -> row, or None
    B: int
    dims: Sequence[int]            # (D, H, W)


def voxmap_words(B, D, H, W):
    n = _lib.load().u3d_voxmap_words(B, D, H, W)
    if n == 0:
        raise _lib.U3DError("B*D*H*W exceeds the 32-bit linear cell index; split the batch")
    return n


def _scan_scratch(words, device):
    n = _lib.load().u3d_scan_scratch_ints(words)
    return torch.empty(n, dtype=torch.int32, device=device)


@dataclass
class Voxels:
    coors: torch.Tensor        # (cap,4) int32 [b,z,y,x]
    feats: torch.Tensor        # (cap,C) f32 VFE mean
    num_points: Optional[torch.Tensor]
    voxels: Optional[torch.Tensor]
    scene_rows: torch.Tensor   # (B+1,) int32 device; [B] = M
    vmap: VoxelMap
    cap: int
    pt_coors: Optional[torch.Tensor] = None  # dynamic mode: (N,4) per-point coords

    @property
    def n_rows(self):
        return self.scene_rows[-1:]


@_timed(lambda r, points, *a, **k: dict(n_points=points.shape[0], C=points.shape[1],
                                        n_voxels=int(r.scene_rows[-1])))
def voxelize_hard(points, pt_off, B, pc_range, voxel_size, grid_zyx, max_pts, max_voxels,
                  deterministic=True, want_voxels=False) -> Voxels:
    lib = _lib.load()
    _req(points, torch.float32, "points")
    _req(pt_off, torch.int32, "pt_off")
    Ntot, C = points.shape
    D, H, W = [int(v) for v in grid_zyx]
    dev = points.device
    words = voxmap_words(B, D, H, W)
    cap = Ntot if max_voxels <= 0 else min(Ntot, B * max_voxels)
    cap = max(cap, 1)
    vm = torch.empty((words, 2), dtype=torch.int32, device=dev)
    scratch = _scan_scratch(words, dev)
    pt_lin = torch.empty(max(Ntot, 1), dtype=torch.int32, device=dev)
    slots = torch.empty(max(Ntot, 1) * max_pts, dtype=torch.int32, device=dev)
    row_of_rank = torch.empty(max(Ntot, 1), dtype=torch.int32, device=dev)
    coors = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    num_points = torch.empty(cap, dtype=torch.int32, device=dev)
    voxels = torch.empty((cap, max_pts, C), dtype=torch.float32, device=dev) if want_voxels else None
    feats = torch.empty((cap, C), dtype=torch.float32, device=dev)
    scene_rows = torch.empty(B + 1, dtype=torch.int32, device=dev)
    pr, prp = _farr(pc_range)
    vs, vsp = _farr(voxel_size)
    order = ORDER_FIRST_APPEARANCE if deterministic else ORDER_LINEAR
    _lib.check(lib.u3d_voxelize_hard(_p(points), _p(pt_off), Ntot, B, C, prp, vsp, D, H, W, max_pts,
                                     int(max_voxels), order, _p(vm), _p(scratch), _p(pt_lin),
                                     _p(slots), _p(row_of_rank), _p(coors), _p(num_points),
                                     _p(voxels), _p(feats), _p(scene_rows), cap, _stream()))
    return Voxels(coors, feats, num_points, voxels, scene_rows,
                  VoxelMap(vm, row_of_rank, B, (D, H, W)), cap)


@_timed(lambda r, points, *a, **k: dict(n_points=points.shape[0], C=points.shape[1],
                                        n_voxels=int(r.scene_rows[-1])))
def voxelize_dynamic(points, pt_off, B, pc_range, voxel_size, grid_zyx) -> Voxels:
    lib = _lib.load()
    _req(points, torch.float32, "points")
    _req(pt_off, torch.int32, "pt_off")
    Ntot, C = points.shape
    D, H, W = [int(v) for v in grid_zyx]
    dev = points.device
    words = voxmap_words(B, D, H, W)
    cap = max(Ntot, 1)
    vm = torch.empty((words, 2), dtype=torch.int32, device=dev)
    scratch = _scan_scratch(words, dev)
    pt_lin = torch.empty(cap, dtype=torch.int32, device=dev)
    pt_coors = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    coors = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    feats = torch.empty((cap, C), dtype=torch.float32, device=dev)
    cnt = torch.empty(cap, dtype=torch.int32, device=dev)
    scene_rows = torch.empty(B + 1, dtype=torch.int32, device=dev)
    pr, prp = _farr(pc_range)
    vs, vsp = _farr(voxel_size)
    _lib.check(lib.u3d_voxelize_dynamic(_p(points), _p(pt_off), Ntot, B, C, prp, vsp, D, H, W,
                                        _p(vm), _p(scratch), _p(pt_lin), _p(pt_coors), _p(coors),
                                        _p(feats), _p(cnt), _p(scene_rows), cap, _stream()))
    return Voxels(coors, feats, cnt, None, scene_rows, VoxelMap(vm, None, B, (D, H, W)), cap,
                  pt_coors=pt_coors)


def voxmap_build(coors, n_rows, cap, B, dims) -> VoxelMap:
    """VoxelMap of an arbitrary (cap,4) int32 coordinate list with a device-side row count."""
    lib = _lib.load()
    _req(coors, torch.int32, "coors")
    D, H, W = [int(v) for v in dims]
    words = voxmap_words(B, D, H, W)
    vm = torch.empty((words, 2), dtype=torch.int32, device=coors.device)
    scratch = _scan_scratch(words, coors.device)
    perm = torch.empty(max(cap, 1), dtype=torch.int32, device=coors.device)
    _lib.check(lib.u3d_voxmap_build(_p(coors), _p(n_rows), cap, B, D, H, W, _p(vm), _p(scratch),
                                    _p(perm), _stream()))
    return VoxelMap(vm, perm, B, (D, H, W))


class Rulebook(torch.Tensor):
    """Neighbour table (27, cap) int32 that also carries the per-128-row tile masks
    (`.tile_mask`, uint32 as int32) the tensor-core conv uses to skip empty kernel offsets."""
    tile_mask = None
    slot_row = None    # sorted rulebooks (rulebook_sort_tiles): output row computed in slot s
    __torch_function__ = torch._C._disabled_torch_function_impl   # ops on it yield plain tensors

    @staticmethod
    def alloc(cap, device):
        # row stride padded to whole 128-row tiles: the conv bulk-copies 512-byte rulebook rows
        pad = max((cap + 127) // 128, 1) * 128
        nbr = torch.empty((27, pad), dtype=torch.int32, device=device)[:, :max(cap, 1)].as_subclass(Rulebook)
        nbr.tile_mask = torch.empty(pad // 128, dtype=torch.int32, device=device)
        return nbr


@_timed(lambda r, coors, n_rows, *a, **k: dict(n_in=int(n_rows), n_out=int(n_rows),
                                               pairs=_live_pairs(r, n_rows)))
def rulebook_subm(coors, n_rows, cap, vmap: VoxelMap, nbr=None):
    lib = _lib.load()
    _req(coors, torch.int32, "coors")
    if nbr is None:
        nbr = Rulebook.alloc(cap, coors.device)
    D, H, W = vmap.dims
    _lib.check(lib.u3d_rulebook_subm(_p(coors), _p(n_rows), cap, _p(vmap.words), _p(vmap.perm),
                                     vmap.B, D, H, W, _p(nbr), nbr.stride(0),
                                     _p(getattr(nbr, "tile_mask", None)), _stream()))
    return nbr


def conv_out_dims(in_dims, stride, pad, k=3):
    return tuple((int(d) + 2 * int(p) - k) // int(s) + 1 for d, s, p in zip(in_dims, stride, pad))


@_timed(lambda r, coors, n_rows, *a, **k: dict(n_in=int(n_rows), n_out=int(r[1]),
                                               pairs=_live_pairs(r[3], r[1])))
def rulebook_down(coors, n_rows, in_cap, vmap: VoxelMap, stride, pad, out_cap=None, sorted_group=None):
    """Strided SparseConv3d geometry: (out_coors, n_out, out VoxelMap, Rulebook, out_cap). sorted_group = None: the
    natural-order table; an int g >= 0: the tile-sorted table built straight from the output coordinates
    (u3d_rulebook_down_sorted; buckets inside groups of g scenes when g > 0), no natural table."""
    lib = _lib.load()
    _req(coors, torch.int32, "coors")
    dev = coors.device
    in_dims = tuple(int(v) for v in vmap.dims)
    out_dims = conv_out_dims(in_dims, stride, pad)
    words = voxmap_words(vmap.B, *out_dims)
    if out_cap is None:
        out_cap = min(8 * in_cap, vmap.B * out_dims[0] * out_dims[1] * out_dims[2])
    out_cap = max(int(out_cap), 1)
    out_vm = torch.empty((words, 2), dtype=torch.int32, device=dev)
    scratch = _scan_scratch(words, dev)
    out_coors = torch.empty((out_cap, 4), dtype=torch.int32, device=dev)
    n_out = torch.empty(1, dtype=torch.int32, device=dev)
    a1, p1 = _iarr(in_dims)
    a2, p2 = _iarr(out_dims)
    a3, p3 = _iarr(stride)
    a4, p4 = _iarr(pad)
    if sorted_group is None:
        nbr = Rulebook.alloc(out_cap, dev)
        _lib.check(lib.u3d_rulebook_down(_p(coors), _p(n_rows), in_cap, _p(vmap.words), _p(vmap.perm),
                                         vmap.B, p1, p2, p3, p4, _p(out_vm), _p(scratch), _p(out_coors),
                                         _p(n_out), out_cap, _p(nbr), nbr.stride(0), _p(nbr.tile_mask),
                                         _stream()))
        return out_coors, n_out, VoxelMap(out_vm, None, vmap.B, out_dims), nbr, out_cap
    _lib.check(lib.u3d_rulebook_down(_p(coors), _p(n_rows), in_cap, _p(vmap.words), _p(vmap.perm),
                                     vmap.B, p1, p2, p3, p4, _p(out_vm), _p(scratch), _p(out_coors),
                                     _p(n_out), out_cap, None, 0, None, _stream()))
    pad256 = (out_cap + 255) // 256 * 256
    srt = torch.empty((27, pad256), dtype=torch.int32, device=dev)[:, :out_cap].as_subclass(Rulebook)
    srt.tile_mask = torch.empty(pad256 // 128, dtype=torch.int32, device=dev)
    srt.slot_row = torch.empty(pad256, dtype=torch.int32, device=dev)
    spg = int(sorted_group)
    G = (int(vmap.B) + spg - 1) // spg if spg else 1
    sscr = torch.empty(lib.u3d_tile_sort_grouped_scratch_ints(out_cap, max(G, 1)), dtype=torch.int32, device=dev)
    _lib.check(lib.u3d_rulebook_down_sorted(_p(out_coors), _p(n_out), out_cap, _p(vmap.words), _p(vmap.perm), vmap.B,
                                            p1, p3, p4, G, spg, _p(sscr), _p(srt.slot_row), _p(srt), srt.stride(0),
                                            _p(srt.tile_mask), _stream()))
    return out_coors, n_out, VoxelMap(out_vm, None, vmap.B, out_dims), srt, out_cap


@_timed(lambda r, nbr, n_out, cap, *a, **k: dict(n_out=int(n_out)))
def rulebook_sort_tiles(nbr, n_out, cap, coors=None, n_scenes=0, scenes_per_group=0):
    """Tile scheduling for the tensor-core conv: returns a SORTED Rulebook (27, cap) whose slots group
    rows with similar neighbour masks (`.slot_row[s]` = output row of slot s, `.tile_mask` per 128
    slots); see csrc/tilesort.cu. Inputs: the natural-order table and its device row count.
    With `coors` (the output rows' (cap,4) coordinates), `n_scenes` and
    `scenes_per_group` > 0 the buckets stay inside groups of consecutive scenes."""
    lib = _lib.load()
    K = nbr.shape[0]
    cap = max(int(cap), 1)
    dev = nbr.device
    pad = (cap + 255) // 256 * 256
    srt = torch.empty((K, pad), dtype=torch.int32, device=dev)[:, :cap].as_subclass(Rulebook)
    srt.tile_mask = torch.empty(pad // 128, dtype=torch.int32, device=dev)
    srt.slot_row = torch.empty(pad, dtype=torch.int32, device=dev)
    if scenes_per_group and coors is not None and n_scenes > 0:
        G = (int(n_scenes) + int(scenes_per_group) - 1) // int(scenes_per_group)
        _req(coors, torch.int32, "coors")
        scratch = torch.empty(lib.u3d_tile_sort_grouped_scratch_ints(cap, G), dtype=torch.int32, device=dev)
        _lib.check(lib.u3d_rulebook_sort_tiles_grouped(_p(nbr), nbr.stride(0), _p(coors), _p(n_out), cap, K, G,
                                                       int(scenes_per_group), _p(scratch), _p(srt.slot_row),
                                                       _p(srt), srt.stride(0), _p(srt.tile_mask), _stream()))
        return srt
    scratch = torch.empty(lib.u3d_tile_sort_scratch_ints(cap), dtype=torch.int32, device=dev)
    _lib.check(lib.u3d_rulebook_sort_tiles(_p(nbr), nbr.stride(0), _p(n_out), cap, K, _p(scratch),
                                           _p(srt.slot_row), _p(srt), srt.stride(0), _p(srt.tile_mask),
                                           _stream()))
    return srt


@_timed(lambda r, coors, n_rows, cap, *a, **k: dict(n_out=int(n_rows)))
def rulebook_subm_sorted(coors, n_rows, cap, vmap: VoxelMap, scenes_per_group=0):
    """The tile-sorted SubM Rulebook (as rulebook_sort_tiles(rulebook_subm(...))) built straight from the coordinates
    and the VoxelMap, without the natural-order table (csrc/tilesort.cu: u3d_rulebook_subm_sorted)."""
    lib = _lib.load()
    _req(coors, torch.int32, "coors")
    cap = max(int(cap), 1)
    dev = coors.device
    pad = (cap + 255) // 256 * 256
    srt = torch.empty((27, pad), dtype=torch.int32, device=dev)[:, :cap].as_subclass(Rulebook)
    srt.tile_mask = torch.empty(pad // 128, dtype=torch.int32, device=dev)
    srt.slot_row = torch.empty(pad, dtype=torch.int32, device=dev)
    G = (int(vmap.B) + int(scenes_per_group) - 1) // int(scenes_per_group) if scenes_per_group else 1
    scratch = torch.empty(lib.u3d_tile_sort_grouped_scratch_ints(cap, max(G, 1)), dtype=torch.int32, device=dev)
    D, H, W = vmap.dims
    _lib.check(lib.u3d_rulebook_subm_sorted(_p(coors), _p(n_rows), cap, _p(vmap.words), _p(vmap.perm), vmap.B, D, H, W,
                                            G, int(scenes_per_group), _p(scratch), _p(srt.slot_row), _p(srt),
                                            srt.stride(0), _p(srt.tile_mask), _stream()))
    return srt


def rulebook_pairs(nbr, n_out):
    """spconv-1.x style (indice_pairs (2,K,N), indice_num (K)) from a neighbour table."""
    lib = _lib.load()
    K, cap = nbr.shape
    pairs = torch.empty((2, K, cap), dtype=torch.int32, device=nbr.device)
    num = torch.empty(K, dtype=torch.int32, device=nbr.device)
    _lib.check(lib.u3d_rulebook_pairs(_p(nbr), nbr.stride(0), _p(n_out), K, _p(pairs[0]),
                                      _p(pairs[1]), cap, _p(num), _stream()))
    return pairs, num


@_timed(lambda r, x, nbr, n_out, out_cap, w, *a, **k: dict(
    n_out=int(n_out), pairs=_live_pairs(nbr, n_out), K=w.shape[0], Cin=w.shape[1], Cout=w.shape[2],
    esize=x.element_size()))
def spconv_fwd(x, nbr, n_out, out_cap, w, scale=None, shift=None, residual=None, relu=False,
               out=None, impl=0):
    """out = act((sum_k x[nbr[k]] @ w[k]) * scale + shift (+ residual)); w is (K,Cin,Cout)."""
    lib = _lib.load()
    _req(x, None, "x")
    dt = _DT[x.dtype]
    K, Cin, Cout = w.shape
    _req(w, x.dtype, "w")
    if out is None:
        out = torch.empty((out_cap, Cout), dtype=x.dtype, device=x.device)
    stride = nbr.stride(0) if nbr is not None else 0
    _lib.check(lib.u3d_spconv_fwd(_p(x), _p(nbr), stride, _p(n_out), out_cap, K, _p(w), _p(scale),
                                  _p(shift), _p(residual), int(bool(relu)), _p(out), Cin, Cout, dt,
                                  impl, _stream()))
    return out


def spconv_tc_supported(K, Cin, Cout):
    return _lib.load().u3d_spconv_packed_bytes(K, Cin, Cout) != 0


def spconv_pack_weights(w):
    """(K,Cin,Cout) bf16 -> packed tensor-core weight image (see include/u3d.h)."""
    lib = _lib.load()
    _req(w, torch.bfloat16, "w")
    K, Cin, Cout = w.shape
    nbytes = lib.u3d_spconv_packed_bytes(K, Cin, Cout)
    if nbytes == 0:
        raise _lib.U3DError(f"tensor-core sparse conv does not support K={K} Cin={Cin} Cout={Cout}")
    packed = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.u3d_spconv_pack_weights(_p(w), K, Cin, Cout, _p(packed), _stream()))
    return packed


@_timed(lambda r, x, nbr, n_out, out_cap, w_packed, K, Cin, Cout, *a, **k: dict(
    n_out=int(n_out), pairs=_live_pairs(nbr, n_out), K=K, Cin=Cin, Cout=Cout, esize=2, tc=True))
def spconv_fwd_packed(x, nbr, n_out, out_cap, w_packed, K, Cin, Cout, scale=None, shift=None,
                      residual=None, relu=False, out=None, reverse=False):
    """tcgen05 sparse conv: bf16 in/out, weights from spconv_pack_weights."""
    lib = _lib.load()
    _req(x, torch.bfloat16, "x")
    if residual is not None:
        _req(residual, torch.bfloat16, "residual")
    if out is None:
        out = torch.empty((out_cap, Cout), dtype=torch.bfloat16, device=x.device)
    stride = nbr.stride(0) if nbr is not None else 0
    tile_mask = getattr(nbr, "tile_mask", None) if nbr is not None else None
    slot_row = getattr(nbr, "slot_row", None) if nbr is not None else None
    _lib.check(lib.u3d_spconv_fwd_packed(_p(x), _p(nbr), stride, _p(tile_mask), _p(slot_row), _p(n_out),
                                         out_cap, K, _p(w_packed),
                                         _p(scale), _p(shift), _p(residual), int(bool(relu)), _p(out),
                                         Cin, Cout, 2 if reverse else 0, _stream()))
    return out


@_timed(lambda r, feats, *a, **k: dict(bytes=r.numel() * r.element_size()))
def sparse_to_dense(feats, coors, n_rows, cap, B, dims, channels_last=True, out=None):
    lib = _lib.load()
    D, H, W = [int(v) for v in dims]
    C = feats.shape[1]
    if out is None:
        shape = (B, D, H, W, C) if channels_last else (B, C, D, H, W)
        out = torch.empty(shape, dtype=feats.dtype, device=feats.device)
    _lib.check(lib.u3d_sparse_to_dense(_p(feats), _p(coors), _p(n_rows), cap, B, D, H, W, C,
                                       _DT[feats.dtype], int(bool(channels_last)), _p(out),
                                       _stream()))
    return out


@_timed(lambda r, dist_src, ds, dss, gs, gst, seg, B, max_n, nq, **k: dict(B=B, max_n=max_n, nq=nq))
def fps(dist_src, dist_stride, dist_seg_stride, gather_src, gather_stride, seg, B, max_n, nq,
        reverse=False, tie_block=1024):
    """Batched D-FPS + gather + min-max normalise. Returns (idx (B,nq) int32, pts (B,nq,3) f32).
    tie_block: exact distance ties resolve like mmcv's kernel with that block-size cap (1024, default) or to the
    lowest index (0)."""
    lib = _lib.load()
    _req(dist_src, torch.float32, "dist_src")
    _req(gather_src, torch.float32, "gather_src")
    _req(seg, torch.int32, "seg")
    idx = torch.empty((B, nq), dtype=torch.int32, device=dist_src.device)
    out = torch.empty((B, nq, 3), dtype=torch.float32, device=dist_src.device)
    _lib.check(lib.u3d_fps(_p(dist_src), dist_stride, dist_seg_stride, _p(gather_src),
                           gather_stride, _p(seg), B, int(max_n), nq, int(bool(reverse)), int(tie_block), _p(idx),
                           _p(out), _stream()))
    return idx, out


def coors_to_float(coors, rows=None):
    lib = _lib.load()
    rows = coors.shape[0] if rows is None else rows
    out = torch.empty((coors.shape[0], 3), dtype=torch.float32, device=coors.device)
    _lib.check(lib.u3d_coors_to_float(_p(coors), rows, _p(out), _stream()))
    return out


@_timed(lambda r, ref, *a, **k: dict(rows=ref.numel() // 3, bytes=r.numel() * r.element_size()))
def sine_embed(ref, dtype=torch.float32):
    lib = _lib.load()
    _req(ref, torch.float32, "ref")
    rows = ref.numel() // 3
    out = torch.empty(ref.shape[:-1] + (384,), dtype=dtype, device=ref.device)
    _lib.check(lib.u3d_sine_embed(_p(ref), rows, _p(out), _DT[dtype], _stream()))
    return out


@_timed(lambda r, a, *x, **k: dict(rows=a.shape[0], bytes=(2 + len([t for t in x[:2] if t is not None]))
                                    * a.numel() * a.element_size()))
def add_layernorm(a, b, c, gamma, beta, eps, relu=False):
    """act(LN(a (+b) (+c)) * gamma + beta) over the last dim; a,b,c (rows, C) same dtype."""
    lib = _lib.load()
    _req(a, None, "a")
    rows, C = a.shape
    for t in (b, c):
        if t is not None:
            _req(t, a.dtype, "residual")
    out = torch.empty_like(a)
    _lib.check(lib.u3d_add_layernorm(_p(a), _p(b), _p(c), _p(gamma), _p(beta), float(eps), rows, C,
                                     int(bool(relu)), _p(out), _DT[a.dtype], _stream()))
    return out


@_timed(lambda r, xs, *a, **k: dict(bytes=(len(xs) + 1) * r.numel() * r.element_size()))
def bias_act_sum(xs, biases, relus):
    """out = sum_i act_i(xs[i] + biases[i]) for up to three same-shape (..., C) tensors (contiguous,
    channels last); biases[i] is an f32 (C,) tensor or None, relus[i] a bool."""
    lib = _lib.load()
    assert 1 <= len(xs) <= 3 and len(biases) == len(xs) == len(relus)
    x0 = xs[0]
    _req(x0, None, "x0")
    for t in xs[1:]:
        _req(t, x0.dtype, "x")
        assert t.shape == x0.shape
    C = x0.shape[-1]
    bs = []
    for b in biases:
        if b is not None:
            _req(b, torch.float32, "bias")
            assert b.numel() == C
        bs.append(b)
    xs = list(xs) + [None] * (3 - len(xs))
    bs = bs + [None] * (3 - len(bs))
    mask = sum(1 << i for i, r in enumerate(relus) if r)
    out = torch.empty_like(x0)
    _lib.check(lib.u3d_bias_act_sum(_p(xs[0]), _p(xs[1]), _p(xs[2]), _p(bs[0]), _p(bs[1]), _p(bs[2]), mask,
                                    x0.numel() // C, C, _DT[x0.dtype], _p(out), _stream()))
    return out


@_timed(lambda r, q, k_, v, n_seq, seq_len, heads: dict(n_seq=n_seq, seq_len=seq_len, heads=heads,
                                                       esize=q.element_size()))
def mha_core(q, k, v, n_seq, seq_len, heads):
    """q,k,v: 2-D views (n_seq*seq_len, heads*32), unit stride in dim 1 (row strides may differ,
    e.g. column slices of a packed QK projection). Returns (n_seq*seq_len, heads*32)."""
    lib = _lib.load()
    for t in (q, k, v):
        if t.stride(1) != 1 or t.dtype != q.dtype or not t.is_cuda:
            raise _lib.U3DError("mha_core: q,k,v must be CUDA, same dtype, unit-stride in dim 1")
    out = torch.empty((n_seq * seq_len, heads * 32), dtype=q.dtype, device=q.device)
    _lib.check(lib.u3d_mha_core(_p(q), _p(k), _p(v), q.stride(0), k.stride(0), v.stride(0), n_seq,
                                seq_len, heads, _p(out), _DT[q.dtype], _stream()))
    return out


@_timed(lambda r, value, ref, query, *a, **k: dict(rows=query.shape[0], C=query.shape[1],
                                                   esize=query.element_size()))
def cross_sample(value_ndhwc, ref, query, query_pos, gate_w, gate_b, Q):
    lib = _lib.load()
    B, D, H, W, C = value_ndhwc.shape
    _req(value_ndhwc, None, "value")
    _req(ref, torch.float32, "ref")
    _req(query, value_ndhwc.dtype, "query")
    out = torch.empty((B * Q, C), dtype=value_ndhwc.dtype, device=value_ndhwc.device)
    _lib.check(lib.u3d_cross_sample(_p(value_ndhwc), B, D, H, W, C, _p(ref), _p(query),
                                    _p(query_pos), _p(gate_w), float(gate_b), Q, _p(out),
                                    _DT[value_ndhwc.dtype], _stream()))
    return out


@_timed(lambda r, boxes, *a, **k: dict(B=boxes.shape[0], N=boxes.shape[1]))
def nms3d_bev(boxes, labels, valid, iou_threshold):
    """boxes (B,N,7) f32 sorted per scene by (label, score desc), labels (B,N) int32, valid (B,N) bool
    -> keep (B,N) bool. Same-label rotated-BEV-IoU greedy NMS for all scenes in one launch pair."""
    lib = _lib.load()
    _req(boxes, torch.float32, "boxes")
    _req(labels, torch.int32, "labels")
    B, N = labels.shape
    v8 = valid.to(torch.uint8).contiguous()
    mask = torch.empty(max(B * lib.u3d_nms3d_mask_words(N), 1), dtype=torch.int64, device=boxes.device)
    keep = torch.empty((B, N), dtype=torch.uint8, device=boxes.device)
    _lib.check(lib.u3d_nms3d_bev(_p(boxes), _p(labels), _p(v8), B, N, float(iou_threshold), _p(mask),
                                 _p(keep), _stream()))
    return keep.bool()


def launch_count():
    """Kernels launched by libu3d_b200 in this process so far."""
    return int(_lib.load().u3d_launch_count())


# ---------------------------------------------------------------- input pre-stage (experimental) ----
def points_prepare(raw, raw_off, B, use_dim, shift_height=False, pc_range=None):
    """Input pre-stage (csrc/points.cu; validated on hardware in round 2). LoadPointsFromFile column select
    (+ shift_height) and PointsRangeFilter on the device. raw (Ntot, load_dim) f32, raw_off (B+1) int32.
    Returns (points (Ntot, C) with the kept rows packed scene by scene, out_off (B+1) int32, floor_z (B))."""
    lib = _lib.load()
    _req(raw, torch.float32, "raw")
    _req(raw_off, torch.int32, "raw_off")
    Ntot, load_dim = raw.shape
    use, usep = _iarr([int(u) for u in use_dim])
    C = len(use_dim) + (1 if shift_height else 0)
    dev = raw.device
    floor_z = torch.empty(B, dtype=torch.float32, device=dev)
    kept = torch.empty(B, dtype=torch.int32, device=dev)
    out = torch.empty((max(Ntot, 1), C), dtype=torch.float32, device=dev)
    out_off = torch.empty(B + 1, dtype=torch.int32, device=dev)
    if pc_range is not None:
        pr, prp = _farr(pc_range)
    else:
        pr, prp = None, None
    _lib.check(lib.u3d_points_prepare(_p(raw), _p(raw_off), B, load_dim, usep, len(use_dim),
                                      int(bool(shift_height)), prp, _p(floor_z), _p(kept), _p(out),
                                      _p(out_off), _stream()))
    return out, out_off, floor_z


def points_gather(points, choices):
    """PointSample with host-drawn indices: points[choices] on the device."""
    lib = _lib.load()
    _req(points, torch.float32, "points")
    _req(choices, torch.int32, "choices")
    n, C = choices.numel(), points.shape[1]
    out = torch.empty((n, C), dtype=torch.float32, device=points.device)
    _lib.check(lib.u3d_points_gather(_p(points), C, _p(choices), n, _p(out), _stream()))
    return out


# ------------------------------------------------------------- decoder / head GEMMs (tcgen05) ----
LIN_RELU1, LIN_MUL, LIN_RES1, LIN_RES2, LIN_LN, LIN_RELU2, LIN_OUT2, LIN_OUT_F32, LIN_REF = \
    1, 2, 4, 8, 16, 32, 64, 128, 256


def linear_supported(N, K):
    return _lib.load().u3d_linear_packed_bytes(int(N), int(K)) != 0


class PackedLinear:
    """nn.Linear parameters in the layout u3d_linear_tc consumes: `w` the pre-swizzled bf16 weight
    images, `bias` f32; `ln` = (gamma f32, beta f32, eps) of a LayerNorm fused behind it, or None."""
    __slots__ = ("w", "bias", "N", "K", "ln")

    def __init__(self, weight, bias=None, ln=None):
        lib = _lib.load()
        N, K = weight.shape
        nbytes = lib.u3d_linear_packed_bytes(N, K)
        if nbytes == 0:
            raise _lib.U3DError(f"u3d_linear_tc does not support N={N} K={K}")
        w = weight.detach().to(torch.bfloat16).contiguous()
        _req(w, torch.bfloat16, "weight")
        self.w = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
        _lib.check(lib.u3d_linear_pack_weights(_p(w), N, K, _p(self.w), _stream()))
        self.bias = None if bias is None else bias.detach().float().contiguous()
        self.N, self.K = N, K
        self.ln = None if ln is None else (ln[0].detach().float().contiguous(), ln[1].detach().float().contiguous(),
                                           float(ln[2]))


@_timed(lambda r, a, lin, **k: dict(rows=a.shape[0], K=lin.K, N=lin.N))
def linear_tc(a, lin, relu=False, mul=None, res1=None, res2=None, ln=False, relu_out=False, add2=None,
              out_f32=False, ref_in=None, out=None, _debug=0):
    """out = act2(LN(act1(a @ W^T + b) * mul + res1 + res2)); returns out, or (out, out + add2) with
    `add2`, or (out f32, ref_in + (out[:,0], out[:,1], out[:,4])) with `ref_in`. a: (rows, K) bf16 2-D
    view with unit column stride; mul / res1 / res2 / add2: (rows, N) bf16 sharing one row stride."""
    lib = _lib.load()
    if a.dtype != torch.bfloat16 or not a.is_cuda or a.dim() != 2 or a.stride(1) != 1:
        raise _lib.U3DError("linear_tc: a must be a CUDA bf16 (rows, K) view with unit column stride")
    rows, K = a.shape
    assert K == lin.K
    N = lin.N
    flags = (LIN_RELU1 if relu else 0) | (LIN_RELU2 if relu_out else 0) | _debug
    ldr = N
    for t, f in ((mul, LIN_MUL), (res1, LIN_RES1), (res2, LIN_RES2), (add2, LIN_OUT2)):
        if t is not None:
            if t.dtype != torch.bfloat16 or t.stride(1) != 1 or t.shape != (rows, N):
                raise _lib.U3DError("linear_tc: epilogue operands must be (rows, N) bf16 with unit column stride")
            flags |= f
            ldr = t.stride(0)
    for t in (mul, res1, res2, add2):
        if t is not None and t.stride(0) != ldr:
            raise _lib.U3DError("linear_tc: epilogue operands must share one row stride")
    g = b = None
    eps = 0.0
    if ln:
        g, b, eps = lin.ln
        flags |= LIN_LN
    out2 = ref_out = None
    if out is not None and (not out.is_contiguous() or tuple(out.shape) != (rows, N) or
                            out.dtype != (torch.float32 if out_f32 else torch.bfloat16)):
        raise _lib.U3DError("linear_tc: `out` must be a contiguous (rows, N) tensor of the output dtype")
    if out_f32:
        flags |= LIN_OUT_F32
        if out is None:
            out = torch.empty((rows, N), dtype=torch.float32, device=a.device)
        if ref_in is not None:
            _req(ref_in, torch.float32, "ref_in")
            flags |= LIN_REF
            ref_out = torch.empty_like(ref_in)
    else:
        if out is None:
            out = torch.empty((rows, N), dtype=torch.bfloat16, device=a.device)
        if add2 is not None:
            out2 = torch.empty((rows, N), dtype=torch.bfloat16, device=a.device)
            if ldr != N:
                raise _lib.U3DError("linear_tc: add2 must be contiguous (out2 shares its row stride)")
    _lib.check(lib.u3d_linear_tc(_p(a), a.stride(0), rows, K, _p(lin.w), N, _p(lin.bias), flags, _p(mul), _p(res1),
                                 _p(res2), ldr, _p(g), _p(b), eps, _p(add2), _p(out2), _p(out), N, _p(ref_in),
                                 _p(ref_out), _stream()))
    if ref_out is not None:
        return out, ref_out
    if out2 is not None:
        return out, out2
    return out


def pos3_ln_relu(ref, weight, bias, gamma, beta, eps, dtype=torch.bfloat16):
    """relu(LayerNorm(ref @ weight^T + bias)) for Linear(3 -> C): ref (rows,3) f32; parameters f32."""
    lib = _lib.load()
    _req(ref, torch.float32, "ref")
    rows, C = ref.shape[0], weight.shape[0]
    out = torch.empty((rows, C), dtype=dtype, device=ref.device)
    _lib.check(lib.u3d_pos3_ln_relu(_p(ref), _p(weight), _p(bias), _p(gamma), _p(beta), float(eps), rows, C,
                                    _p(out), _DT[dtype], _stream()))
    return out


def box_assemble(tmp, ref_logit, pc_range, out=None):
    """Uni3DETRHead.forward's box assembly (uni3detr_head.py:470-496) for (rows, code) f32 raw branch
    outputs and the (rows, 3) f32 reference LOGITS the layer started from."""
    lib = _lib.load()
    _req(tmp, torch.float32, "tmp")
    _req(ref_logit, torch.float32, "ref_logit")
    rows, code = tmp.shape
    if out is None:
        out = torch.empty_like(tmp)
    pr, prp = _farr(pc_range)
    _lib.check(lib.u3d_box_assemble(_p(tmp), _p(ref_logit), rows, code, prp, _p(out), _stream()))
    return out


def is_dense(x):
    """Contiguous in the default or the channels-last layout of its rank (elementwise kernels may then
    walk the storage linearly)."""
    if x.is_contiguous():
        return True
    if x.dim() == 4:
        return x.is_contiguous(memory_format=torch.channels_last)
    if x.dim() == 5:
        return x.is_contiguous(memory_format=torch.channels_last_3d)
    return False


def split_tf32(x):
    """(hi, lo) with hi = x rounded to TF32 and lo = x - hi (both fp32, same shape / strides as x)."""
    lib = _lib.load()
    if x.dtype != torch.float32 or not x.is_cuda:
        raise _lib.U3DError("split_tf32: x must be a CUDA fp32 tensor")
    hi, lo = torch.empty_like(x), torch.empty_like(x)     # preserve_format: same memory layout as x
    if not (is_dense(x) and hi.stride() == x.stride()):
        raise _lib.U3DError("split_tf32: x must be dense (contiguous or channels-last)")
    _lib.check(lib.u3d_split_tf32(_p(x), x.numel(), _p(hi), _p(lo), _stream()))
    return hi, lo


# ------------------------------------------------------------------- training-side kernels ----
def rulebook_transpose(nbr, n_out, out_cap, in_cap):
    """(K, in_cap) table of the data gradient: nbr_t[k][i] = o iff nbr[k][o] = i (else -1)."""
    lib = _lib.load()
    K = nbr.shape[0]
    pad = max((int(in_cap) + 127) // 128, 1) * 128
    nbr_t = torch.empty((K, pad), dtype=torch.int32, device=nbr.device)[:, :max(int(in_cap), 1)]
    _lib.check(lib.u3d_rulebook_transpose(_p(nbr), nbr.stride(0), _p(n_out), int(out_cap), K, _p(nbr_t),
                                          nbr_t.stride(0), int(in_cap), _stream()))
    return nbr_t


def spconv_wgrad(x, dy, nbr, n_out, out_cap, K, Cin, Cout):
    """dW (K,Cin,Cout) f32 = sum over rulebook pairs of x[in]^T dy[out]."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(dy, torch.float32, "dy")
    dW = torch.empty((K, Cin, Cout), dtype=torch.float32, device=x.device)
    _lib.check(lib.u3d_spconv_wgrad(_p(x), _p(dy), _p(nbr), nbr.stride(0), _p(n_out), int(out_cap), K, Cin, Cout,
                                    _p(dW), _stream()))
    return dW


def cross_sample_bwd(value_ndhwc, ref, query, query_pos, gate_w, gate_b, Q, d_out, need_value_grad=True):
    """Backward of cross_sample (fp32). Returns (d_value or None, d_q, d_gate_w, d_gate_b, d_ref)."""
    lib = _lib.load()
    B, D, H, W, C = value_ndhwc.shape
    for t, n in ((value_ndhwc, "value"), (ref, "ref"), (query, "query"), (d_out, "d_out"), (gate_w, "gate_w")):
        _req(t, torch.float32, n)
    dev = query.device
    d_value = torch.empty_like(value_ndhwc) if need_value_grad else None
    d_q = torch.empty_like(query)
    d_gw = torch.empty(C, dtype=torch.float32, device=dev)
    d_gb = torch.empty(1, dtype=torch.float32, device=dev)
    d_ref = torch.empty_like(ref)
    _lib.check(lib.u3d_cross_sample_bwd(_p(value_ndhwc), B, D, H, W, C, _p(ref), _p(query), _p(query_pos), _p(gate_w),
                                        float(gate_b), Q, _p(d_out), _p(d_value), _p(d_q), _p(d_gw), _p(d_gb),
                                        _p(d_ref), _stream()))
    return d_value, d_q, d_gw, d_gb, d_ref


def iou3d_aligned(a, b):
    """Rotated 3-D IoU of box pairs: a, b (n,7) f32 [x,y,z(bottom),dx,dy,dz,yaw] -> (n,) f32."""
    lib = _lib.load()
    _req(a, torch.float32, "a")
    _req(b, torch.float32, "b")
    out = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
    _lib.check(lib.u3d_iou3d_aligned(_p(a), _p(b), a.shape[0], _p(out), _stream()))
    return out


def hungarian(cost):
    """cost (P, rows, cols) f32 CUDA, rows <= cols -> (P, rows) int32: the column assigned to every row by the
    minimum-cost assignment (what scipy.optimize.linear_sum_assignment(cost[p]) returns as col_ind)."""
    lib = _lib.load()
    _req(cost, torch.float32, "cost")
    P, rows, cols = cost.shape
    out = torch.empty((P, rows), dtype=torch.int32, device=cost.device)
    _lib.check(lib.u3d_hungarian(_p(cost), rows * cols, cols, P, rows, cols, _p(out), _stream()))
    return out


# ------------------------------------------------------------- fp32 sparse conv as 3xBF16 (tcgen05) ----
def split_bf16(x):
    """fp32 (rows, C) -> (rows, 2C) bf16 [hi | lo] with hi = bf16(x), lo = bf16(x - hi)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=1).contiguous()


def merge_bf16(x2):
    """(rows, 2C) bf16 [hi | lo] -> fp32 (rows, C)."""
    C = x2.shape[1] // 2
    return x2[:, :C].float() + x2[:, C:].float()


class PackedConvX3:
    """Weights of one fp32 sparse conv for u3d_spconv_fwd_packed_x3: per <=128-channel output part the packed
    images of the K-block groups [w_hi ; w_lo ; w_hi] (built from two u3d_spconv_pack_weights calls)."""

    def __init__(self, w):
        K, Cin, Cout = w.shape
        self.K, self.Cin, self.Cout = K, Cin, Cout
        w = w.float()
        w_hi = w.to(torch.bfloat16)
        w_lo = (w - w_hi.float()).to(torch.bfloat16)
        self.parts = []
        for c0 in range(0, Cout, 128):
            c1 = min(c0 + 128, Cout)
            co = c1 - c0
            ph = spconv_pack_weights(w_hi[:, :, c0:c1].contiguous())
            pl = spconv_pack_weights(w_lo[:, :, c0:c1].contiguous())
            n1 = K * Cin * co                      # elements of the first image set (rows-on-M, Cout rows per image)
            sets = [(0, n1, co)]
            if co < 128:
                sets.append((n1, n1 + K * Cin * 128, 128))     # replicated 128-row images of the rows-on-N kernel
            blk = 64 if Cin % 64 == 0 else Cin
            nkb = Cin // blk
            chunks = []
            for a, b, rows in sets:
                vh = ph[a:b].view(K, nkb, rows * blk)
                vl = pl[a:b].view(K, nkb, rows * blk)
                chunks.append(torch.cat([vh, vl, vh], dim=1).reshape(-1))
            self.parts.append((c0, co, torch.cat(chunks).contiguous()))


def identity_rulebook(cap, device):
    """(1, cap) table mapping every output row to itself: lets 1x1x1 convs run on the gather kernel."""
    pad = max((cap + 255) // 256, 1) * 256
    nbr = torch.arange(pad, dtype=torch.int32, device=device).view(1, pad)[:, :max(cap, 1)].as_subclass(Rulebook)
    nbr.tile_mask = None
    return nbr


@_timed(lambda r, x2, nbr, n_out, out_cap, pk, *a, **k: dict(
    n_out=int(n_out), pairs=_live_pairs(nbr, n_out), K=pk.K, Cin=pk.Cin, Cout=pk.Cout, esize=4, tc=True, x3=True))
def spconv_fwd_packed_x3(x2, nbr, n_out, out_cap, pk, scale=None, shift=None, residual=None, relu=False, reverse=False):
    """fp32-grade sparse conv on tcgen05: x2 (rows, 2*Cin) bf16 [hi|lo] -> (out_cap, 2*Cout) bf16 [hi|lo]."""
    lib = _lib.load()
    _req(x2, torch.bfloat16, "x2")
    assert x2.shape[1] == 2 * pk.Cin
    out = torch.empty((out_cap, 2 * pk.Cout), dtype=torch.bfloat16, device=x2.device)
    tile_mask = getattr(nbr, "tile_mask", None)
    slot_row = getattr(nbr, "slot_row", None)
    for c0, co, wp in pk.parts:
        _lib.check(lib.u3d_spconv_fwd_packed_x3(
            _p(x2), _p(nbr), nbr.stride(0), _p(tile_mask), _p(slot_row), _p(n_out), out_cap, pk.K, _p(wp),
            _p(scale[c0:c0 + co]) if scale is not None else None, _p(shift[c0:c0 + co]) if shift is not None else None,
            _p(residual), int(bool(relu)), _p(out), pk.Cin, co, c0, pk.Cout, 2 if reverse else 0, _stream()))
    return out
