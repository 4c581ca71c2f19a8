"""ctypes binding of libgonzag_b200.so (the C ABI declared in include/gonzag_b200.h).

There is no Python or CPU fallback: if the shared library is missing or no CUDA device is
present, the product raises `GonzagB200Error`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgonzag_b200.so")

GZB_OK, GZB_ERR_CUDA, GZB_ERR_ARG, GZB_ERR_NOMEM, GZB_ERR_STATE, GZB_ERR_TIME = range(6)
GZB_HOST, GZB_DEVICE = 0, 1
GZB_F32, GZB_F64 = 0, 1


class GonzagB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gonzag_b200 error %d: %s" % (code, msg))
        self.code = code


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("last_breaks", C.c_int64), ("last_chain_steps", C.c_int64),
                ("pyramid_levels", C.c_int32), ("l2_fetch_bytes", C.c_int32),
                ("lut_bins", C.c_int64), ("last_unresolved", C.c_int64),
                ("last_unres_noseed", C.c_int64), ("last_unres_noconv", C.c_int64),
                ("last_unres_nocert", C.c_int64), ("last_max_chain", C.c_int64),
                ("heading_table", C.c_int64)]


_P = C.c_void_p
_I64 = C.c_int64
_SIGNATURES = {
    "gzb_version": (C.c_char_p, []),
    "gzb_last_global_error": (C.c_char_p, []),
    "gzb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gzb_create": (C.c_int, [C.c_int, _I64, _I64, _P, _P, _P, _P, C.c_int, C.POINTER(_P)]),
    "gzb_create_imask": (C.c_int, [C.c_int, _I64, _I64, _P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(_P)]),
    "gzb_destroy": (C.c_int, [_P]),
    "gzb_last_error": (C.c_char_p, [_P]),
    "gzb_set_stream": (C.c_int, [_P, _P]),
    "gzb_use_own_stream": (C.c_int, [_P]),
    "gzb_sync": (C.c_int, [_P]),
    "gzb_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "gzb_set_mask": (C.c_int, [_P, _P, C.c_int]),
    "gzb_set_angle": (C.c_int, [_P, _P, C.c_int]),
    "gzb_set_ew_per": (C.c_int, [_P, C.c_int]),
    "gzb_set_option": (C.c_int, [_P, C.c_char_p, _I64]),
    "gzb_grid_angle": (C.c_int, [_P, _P, C.c_int]),
    "gzb_get_angle": (C.c_int, [_P, _P, C.c_int]),
    "gzb_nearest": (C.c_int, [_P, _I64, _P, _P, C.c_double, C.c_int, _I64, _I64, _P, C.c_int]),
    "gzb_source_mesh": (C.c_int, [_P, _I64, _P, _P, _P, _P, C.c_int]),
    "gzb_weights": (C.c_int, [_P, _I64, _P, _P, _P, _P, C.c_int]),
    "gzb_mesh_weights": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, C.c_int]),
    "gzb_mapping": (C.c_int, [_P, _I64, _P, _P, C.c_double, C.c_int, _I64, _I64, _P, _P, _P, C.c_int]),
    "gzb_lon_stats": (C.c_int, [_P, _P, _P]),
    "gzb_lon_to_pm180": (C.c_int, [_P, _P]),
    "gzb_lat_range": (C.c_int, [_P, _P, _P]),
    "gzb_map_interp": (C.c_int, [_P, _I64, _P, _P, _P, C.c_double, C.c_int, _I64, _I64, _I64, _P, _P, C.c_int, _I64, _I64,
                                 C.c_double, _P, _P, _P, _P, _P]),
    "gzb_map_interp_multi": (C.c_int, [_P, C.c_int, _P, _P, _I64, _P, _P, _P, C.c_double, C.c_int, _I64, _P, _P, C.c_int, _I64,
                                       _I64, C.c_double, _P, _P, _P, _P, _P]),
    "gzb_mesh_weights_interp": (C.c_int, [_P, _I64, _P, _P, _P, _P, _I64, _P, _P, C.c_int, _I64, _I64, C.c_double,
                                          _P, _P, _P, _P]),
    "gzb_interp": (C.c_int, [_P, _I64, _P, _P, _P, _I64, _P, _P, C.c_int, _I64, _I64, C.c_double, C.c_int,
                             _P, _P, C.c_int]),
    "gzb_track_distance": (C.c_int, [_P, _I64, _P, _P, _P, C.c_int]),
    "gzb_xnp_track": (C.c_int, [_P, _I64, _P, C.c_double, _P, C.c_int]),
    "gzb_xnp_track_masked": (C.c_int, [_P, _I64, _P, C.c_double, _P, _P, C.c_int]),
    "gzb_heading": (C.c_int, [C.c_int, _I64, _P, _P, _P, _P, _P]),
    "gzb_alfabeta": (C.c_int, [C.c_int, _I64, _P, _P, _P]),
    "gzb_haversine": (C.c_int, [C.c_int, _I64, C.c_double, C.c_double, _P, _P, _P]),
    "gzb_haversine_sclr": (C.c_int, [C.c_int, _I64, _P, _P, _P, _P, _P]),
    "gzb_radius_earth": (C.c_int, [C.c_int, _I64, _P, _P]),
    "gzb_dev_alloc": (C.c_int, [_P, C.c_uint64, C.POINTER(_P)]),
    "gzb_dev_free": (C.c_int, [_P, _P]),
    "gzb_h2d": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "gzb_d2h": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "gzb_d2d": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "gzb_memset": (C.c_int, [_P, _P, C.c_int, C.c_uint64]),
    "gzb_ipc_export": (C.c_int, [_P, _P, _P]),
    "gzb_ipc_open": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "gzb_ipc_close": (C.c_int, [_P, _P]),
    "gzb_set_peer_outputs": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "gzb_publish_edges": (C.c_int, [_P, _P, _I64, C.c_int, _P]),
    "gzb_pool_config": (C.c_int, [C.c_int, C.c_uint64]),
    "gzb_pool_trim": (C.c_int, []),
    "gzb_pool_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gzb_peer_barrier": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _I64]),
    "gzb_peer_finish": (C.c_int, [_P, _P, _I64, C.c_int, _P, C.c_int, _P, _P, C.c_int, C.c_int, _I64, _P, _I64, _I64, _P]),
    "gzb_check_edges": (C.c_int, [_P, _P, _I64, _I64, C.c_int, _P]),
    "gzb_pack_mask": (C.c_int, [_P, C.c_int, _I64, _P]),
    "gzb_track_mask": (C.c_int, [_P, _I64, _P, _P, _P, _P]),
    "gzb_set_host_threads": (C.c_int, [C.c_int]),
    "gzb_grid_arrays": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "gzb_copy_swap": (C.c_int, [_P, _P, C.c_int, _I64, C.c_int]),
    "gzb_pointer_kind": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "gzb_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(_P)]),
    "gzb_host_free": (C.c_int, [_P]),
    "gzb_host_register": (C.c_int, [_P, C.c_uint64]),
    "gzb_host_unregister": (C.c_int, [_P]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GonzagB200Error(GZB_ERR_STATE,
                                  "%s is missing: build it with `python -m gonzag_b200.build` "
                                  "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(code, ctx=None):
    if code != GZB_OK:
        L = lib()
        msg = (L.gzb_last_error(ctx) if ctx else L.gzb_last_global_error()) or b""
        raise GonzagB200Error(code, msg.decode("utf-8", "replace"))


def ptr(a):
    """Raw pointer of a C-contiguous NumPy array (or None, or an int device pointer)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return C.c_void_p(a.ctypes.data)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def pointer_kind(arr_or_ptr):
    """0 pageable host, 1 pinned + mapped host, 2 device memory."""
    k = C.c_int(0)
    check(lib().gzb_pointer_kind(ptr(arr_or_ptr), C.byref(k)))
    return k.value


def device_count():
    n = C.c_int(0)
    rc = lib().gzb_device_count(C.byref(n))
    return n.value if rc == GZB_OK else 0


def pinned_empty(shape, dtype):
    """NumPy array backed by pinned (page-locked, mapped) host memory; keeps itself alive
    through a finalizer on the base object."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib().gzb_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[p.value] = buf
    return arr


def pinned_free(arr):
    addr = arr.ctypes.data
    if addr in _PINNED:
        del _PINNED[addr]
        lib().gzb_host_free(C.c_void_p(addr))


_PINNED = {}
