"""Build + ctypes binding of libu3d_b200.so (the C ABI declared in include/u3d.h).

The library is compiled in-tree with nvcc for sm_100a only; there is no CPU or
PyTorch fallback: if the shared object is missing and cannot be built, importing
an op raises.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libu3d_b200.so")
_SOURCES = ["voxmap.cu", "voxelize.cu", "rulebook.cu", "spconv_simt.cu", "spconv_tc.cu", "spconv_tn.cu", "tilesort.cu", "points.cu",
            "fps.cu", "decoder.cu", "mha_tc.cu", "mha_tc2.cu", "linear_tc.cu", "train.cu", "nms.cu"]
_HEADERS = [os.path.join(_CSRC, "common.cuh"), os.path.join(_CSRC, "tc_common.cuh"), os.path.join(_CSRC, "bev_geom.cuh"), os.path.join(_CSRC, "tma.cuh"),
            os.path.join(_HERE, "..", "include", "u3d.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]

_lock = threading.Lock()
_lib = None


def so_path():
    return _SO


def source_hash():
    """sha256 over every CUDA source, header and the compiler flags: the library's build id. It is
    compiled into the .so (u3d_build_id()), so a binary is recognised as stale - or as built from a
    different ABI - by content, not by file times (snapshots onto a GPU box do not keep mtimes)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in [os.path.join(_CSRC, s) for s in _SOURCES] + _HEADERS:
        with open(d, "rb") as f:
            h.update(os.path.basename(d).encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def _so_build_id():
    """Build id of the library on disk, read from the file (no dlopen): the marker string
    'U3D_BUILD_ID=<hash>' that csrc/voxmap.cu embeds."""
    if not os.path.exists(_SO):
        return None
    with open(_SO, "rb") as f:
        blob = f.read()
    i = blob.find(b"U3D_BUILD_ID=")
    if i < 0:
        return None
    return blob[i + 13:i + 29].decode("ascii", "replace")


def _stale():
    return _so_build_id() != source_hash()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into libu3d_b200.so (nvcc cross-compiles on CPU).
    Serialised across processes (one rank per GPU may import at the same time): the first one
    builds into a temporary file and renames it, the others find a fresh library."""
    if not force and not _stale():
        return _SO
    import fcntl
    with open(_SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return _SO
            nvcc = os.environ.get("NVCC", "nvcc")
            tmp = _SO + ".tmp.%d" % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + ["-DU3D_BUILD_ID_STR=\"%s\"" % source_hash()] + \
                (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + \
                [os.path.join(_CSRC, s) for s in _SOURCES]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, _SO)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


_vp, _i32, _f32, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/u3d.h declares
SIGNATURES = {
    "u3d_last_error": (ctypes.c_char_p, []),
    "u3d_version": (_i32, []),
    "u3d_build_id": (ctypes.c_char_p, []),
    "u3d_launch_count": (ctypes.c_ulonglong, []),
    "u3d_voxmap_words": (_sz, [_i32] * 4),
    "u3d_scan_scratch_ints": (_sz, [_sz]),
    "u3d_voxelize_hard": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32,
                                 _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "u3d_voxelize_dynamic": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "u3d_voxmap_build": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "u3d_rulebook_subm": (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp]),
    "u3d_rulebook_down": (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _i32, _vp, _i32, _vp, _vp]),
    "u3d_rulebook_pairs": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _vp]),
    "u3d_spconv_fwd": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i32,
                              _i32, _i32, _i32, _vp]),
    "u3d_spconv_packed_bytes": (_sz, [_i32, _i32, _i32]),
    "u3d_spconv_pack_weights": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "u3d_spconv_fwd_packed": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp,
                                     _i32, _vp, _i32, _i32, _i32, _vp]),
    "u3d_spconv_fwd_packed_x3": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp,
                                        _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "u3d_tile_sort_scratch_ints": (_sz, [_i32]),
    "u3d_tile_sort_grouped_scratch_ints": (_sz, [_i32, _i32]),
    "u3d_rulebook_subm_sorted": (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32,
                                 _vp, _vp]),
    "u3d_rulebook_down_sorted": (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32,
                                 _vp, _vp]),
    "u3d_rulebook_sort_tiles_grouped": (_i32, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32,
                                               _vp, _vp]),
    "u3d_rulebook_sort_tiles": (_i32, [_vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "u3d_sparse_to_dense": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                   _vp, _vp]),
    "u3d_fps": (_i32, [_vp, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "u3d_coors_to_float": (_i32, [_vp, _i32, _vp, _vp]),
    "u3d_sine_embed": (_i32, [_vp, _i32, _vp, _i32, _vp]),
    "u3d_add_layernorm": (_i32, [_vp, _vp, _vp, _vp, _vp, _f32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "u3d_points_prepare": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "u3d_points_gather": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "u3d_bias_act_sum": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, ctypes.c_longlong, _i32, _i32, _vp, _vp]),
    "u3d_mha_core": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "u3d_linear_packed_bytes": (_sz, [_i32, _i32]),
    "u3d_linear_pack_weights": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "u3d_linear_tc": (_i32, [_vp, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _f32,
                             _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "u3d_rulebook_transpose": (_i32, [_vp, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _vp]),
    "u3d_spconv_wgrad": (_i32, [_vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "u3d_cross_sample_bwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _f32, _i32, _vp,
                                    _vp, _vp, _vp, _vp, _vp, _vp]),
    "u3d_iou3d_aligned": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "u3d_hungarian_smem_bytes": (_sz, [_i32, _i32]),
    "u3d_hungarian": (_i32, [_vp, ctypes.c_longlong, _i32, _i32, _i32, _i32, _vp, _vp]),
    "u3d_split_tf32": (_i32, [_vp, ctypes.c_longlong, _vp, _vp, _vp]),
    "u3d_box_assemble": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "u3d_pos3_ln_relu": (_i32, [_vp, _vp, _vp, _vp, _vp, _f32, _i32, _i32, _vp, _i32, _vp]),
    "u3d_nms3d_mask_words": (_sz, [_i32]),
    "u3d_nms3d_bev": (_i32, [_vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp]),
    "u3d_cross_sample": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _f32, _i32,
                                _vp, _i32, _vp]),
}


def load():
    """Return the ctypes handle; builds the library first if the sources are newer."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            # no silent fallback and no stale binary: a library built from other sources may export the
            # same names with different argument lists, so a failed rebuild is fatal even if an old
            # .so is still on disk
            try:
                build()
            except Exception as e:
                raise RuntimeError(
                    "libu3d_b200.so is missing or was built from different sources (build id "
                    f"{_so_build_id()} != {source_hash()}) and could not be rebuilt; "
                    "run `python -c 'import __graft_entry__ as g; g.build()'`") from e
        lib = ctypes.CDLL(_SO)
        lib.u3d_build_id.restype = ctypes.c_char_p
        bid = lib.u3d_build_id().decode()
        if bid != "U3D_BUILD_ID=" + source_hash():
            raise RuntimeError(f"libu3d_b200.so build id {bid} does not match the sources ({source_hash()})")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and library diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


class U3DError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().u3d_last_error().decode("utf-8", "replace")
        raise U3DError(f"libu3d_b200 error {rc}: {msg}")
