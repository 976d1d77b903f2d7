"""ctypes binding of libufv_b200.so (the C ABI declared in include/ufv_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C ufvideo_b200/csrc``.
There is no fallback: if the shared object is missing, loading raises and every operator of
this package fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UFV_B200_LIB") or os.path.join(HERE, "libufv_b200.so")   # env: developer sweeps
CSRC = os.path.join(HERE, "csrc")
HEADER = os.path.join(os.path.dirname(HERE), "include", "ufv_b200.h")

UFV_F32, UFV_BF16, UFV_F16, UFV_U8, UFV_RLE = 0, 1, 2, 3, 4
BITS_WORDS = 24
MAX_PATCH_SIDE = 27
MAX_GROUP = 64
ABI_VERSION = 9

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64


MAX_PEER_DST = 8


class PeerArgs(C.Structure):
    """Mirror of ``struct ufv_peer_args`` (include/ufv_b200.h)."""
    _fields_ = [
        ("dst", C.c_uint64 * MAX_PEER_DST), ("tail_dst", C.c_uint64 * MAX_PEER_DST),
        ("flag", C.c_uint64 * MAX_PEER_DST), ("tail_src", _p), ("ticket", _p),
        ("tail_words", _i32), ("n_dst", _i32), ("multimem", _i32), ("flag_value", _i32),
    ]


class EncodeArgs(C.Structure):
    """Mirror of ``struct ufv_encode_args`` (include/ufv_b200.h)."""
    _fields_ = [
        ("feats", _p), ("feat_dtype", _i32), ("n_patch_side", _i32), ("n_rows", _i64),
        ("c", _i32), ("hid", _i32),
        ("mask_desc", _p), ("taps", _p), ("n_masks", _i32), ("idx_pitch", _i32),
        ("any_row_mode", _i32), ("reserved1", _i32),
        ("bits", _p), ("cnt", _p), ("idx", _p),
        ("grp_row", _p), ("grp_off", _p), ("grp_member", _p), ("n_groups", _i32),
        ("max_group", _i32),
        ("pooled", _p),
        ("obj_start", _p), ("obj_len", _p), ("slot_off", _p),
        ("n_obj", _i32), ("max_len", _i32), ("k_keep", _i32), ("m_pad", _i32),
        ("merged", _p), ("counts", _p), ("sims", _p), ("sims_pitch", _i32), ("reserved0", _i32),
        ("counts_host", _p), ("epoch", _i32), ("reserved", _i32),
        ("w1", _p), ("b1", _p), ("w2", _p), ("b2", _p),
        ("hidden", _p), ("tokens_out", _p), ("gemm_ws", _p), ("gemm_ws_bytes", _i64), ("tokens_row_map", _p),
        ("peer", C.POINTER(PeerArgs)),
        ("dyn_src", _p), ("dyn_dev", _p),
    ]


class DynArgs(C.Structure):
    """Mirror of ``struct ufv_dyn_args`` (include/ufv_b200.h): the per-call block of a replayed graph."""
    _fields_ = [("tokens_out", C.c_uint64), ("counts_out", C.c_uint64), ("epoch", _i32), ("reserved", _i32),
                ("peer", PeerArgs), ("pad", C.c_uint64)]


assert C.sizeof(DynArgs) == 256


_SIGNATURES = {
    "ufv_abi_version": (C.c_int, []),
    "ufv_last_error": (C.c_char_p, []),
    "ufv_struct_size": (C.c_int, [C.c_char_p]),
    "ufv_device_address": (C.c_int, [_p, C.POINTER(C.c_uint64)]),
    "ufv_tap_table": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "ufv_mask_to_patches": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p, C.c_int, _p]),
    "ufv_mask_pool": (C.c_int, [_p, C.c_int, _i64, C.c_int, C.c_int, _p, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p]),
    "ufv_mask_pool_backward": (C.c_int, [_p, _p, _p, _p, _i64, C.c_int, C.c_int, C.c_int, _p, C.c_int, _p]),
    "ufv_ttm": (C.c_int, [_p, C.c_int, _p, _p, _p, C.c_int, C.c_int, C.c_int, _p, C.c_int, _p, _p,
                          _p, C.c_int, _p, C.c_int, _p, _i32, _p]),
    "ufv_linear_ws_bytes": (_i64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ufv_linear": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p, _i64, _p]),
    "ufv_linear_ex": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p]),
    "ufv_transpose16": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _p]),
    "ufv_colsum": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _p]),
    "ufv_linear_scatter": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _i64, _p]),
    "ufv_splice_static": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_int, C.c_int, _i64, _p]),
    "ufv_linear_gather": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(PeerArgs), _p, _i64,
                                    _p]),
    "ufv_peer_push": (C.c_int, [_p, _i64, C.POINTER(PeerArgs), _p]),
    "ufv_wait_flags": (C.c_int, [_p, C.c_int, _i32, C.c_int, _p, _p]),
    "ufv_encode_graph_create": (C.c_int, [C.POINTER(EncodeArgs), C.POINTER(_p)]),
    "ufv_encode_graph_launch": (C.c_int, [_p, _p]),
    "ufv_encode_graph_destroy": (C.c_int, [_p]),
    "ufv_encode": (C.c_int, [C.POINTER(EncodeArgs), _p]),
    "ufv_compact_rows": (C.c_int, [_p, _p, _p, C.c_int, _p, C.c_int, _p]),
    "ufv_splice_rows": (C.c_int, [_p, C.c_int, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p, _p, C.c_int, _p]),
    "ufv_gather_rows": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, _p]),
}
EXPORTED = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


def build(verbose: bool = False) -> str:
    """Compile libufv_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libufv_b200.so failed (see output above)")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.isfile(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `make -C ufvideo_b200/csrc` or "
                        "`python -c 'import __graft_entry__ as g; g.build()'`. ufvideo_b200 has no "
                        "CPU or PyTorch fallback for the object-encoder path.")
                handle = C.CDLL(LIB_PATH)
                for name, (restype, argtypes) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = restype
                    fn.argtypes = argtypes
                got = handle.ufv_abi_version()
                if got != ABI_VERSION:
                    raise RuntimeError(f"libufv_b200.so ABI {got} != binding ABI {ABI_VERSION}: rebuild")
                _lib = handle
    return _lib


class UfvError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libufv_b200 error {code}: {message}")
        self.code = code


def check(code: int) -> None:
    if code != 0:
        raise UfvError(code, lib().ufv_last_error().decode("utf-8", "replace"))
