"""ctypes binding of the C-ABI in include/deltarice_b200.h (libh5deltarice_b200.so).

The library holds the sm_100a kernels, the chunk scheduler and the HDF5 filter boundary.
There is no Python/CPU implementation behind it: if the shared library is missing this
module raises (build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C deltarice_b200/csrc`)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libh5deltarice_b200.so")

OK, E_PARAM, E_CUDA, E_CAPACITY, E_STREAM, E_NOMEM, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
_ERRNAMES = {E_PARAM: "DRICE_E_PARAM", E_CUDA: "DRICE_E_CUDA", E_CAPACITY: "DRICE_E_CAPACITY",
             E_STREAM: "DRICE_E_STREAM", E_NOMEM: "DRICE_E_NOMEM", E_UNSUPPORTED: "DRICE_E_UNSUPPORTED"}


class DeltaRiceError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("M", C.c_int32), ("L", C.c_int32), ("filter_len", C.c_int32), ("filter", C.c_int32 * 16)]


# every symbol include/deltarice_b200.h and include/deltaRice.h declare
C_ABI_SYMBOLS = [
    "drice_abi_version", "drice_log2_param", "drice_parse_cd_values", "drice_chunk_bound_bytes",
    "drice_batch_bound_bytes", "drice_create", "drice_destroy", "drice_last_error", "drice_device",
    "drice_set_filter", "drice_host_alloc", "drice_host_free", "drice_encode_batch_dev_async", "drice_encode_batch_dev",
    "drice_decode_batch_dev_async", "drice_decode_batch_dev", "drice_encode_batch_host",
    "drice_decode_batch_host", "drice_peek_chunk_samples", "drice_launch_count",
    "drice_timing_enable", "drice_timing_read", "drice_kernel_name",
]
H5_SYMBOLS = ["H5Z_DELTARICE", "H5Z_filter_deltarice", "deltarice_register_h5filter",
              "H5PLget_plugin_type", "H5PLget_plugin_info"]

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library is not built and deltarice_b200 has no "
            "CPU fallback. Run `make -C deltarice_b200/csrc` (or __graft_entry__.build()).")
    L = C.CDLL(LIB_PATH)
    vp, sz, i, i64, u64p = C.c_void_p, C.c_size_t, C.c_int, C.c_int64, C.POINTER(C.c_uint64)
    L.drice_abi_version.restype = i
    L.drice_log2_param.restype = i
    L.drice_log2_param.argtypes = [i]
    L.drice_parse_cd_values.restype = i
    L.drice_parse_cd_values.argtypes = [sz, C.POINTER(C.c_uint), C.POINTER(Params)]
    L.drice_chunk_bound_bytes.restype = sz
    L.drice_chunk_bound_bytes.argtypes = [sz, i64]
    L.drice_batch_bound_bytes.restype = sz
    L.drice_batch_bound_bytes.argtypes = [u64p, sz, i64]
    L.drice_create.restype = i
    L.drice_create.argtypes = [C.POINTER(vp), i]
    L.drice_destroy.restype = None
    L.drice_destroy.argtypes = [vp]
    L.drice_last_error.restype = C.c_char_p
    L.drice_last_error.argtypes = [vp]
    L.drice_device.restype = i
    L.drice_device.argtypes = [vp]
    L.drice_set_filter.restype = i
    L.drice_set_filter.argtypes = [vp, C.POINTER(C.c_int32), i]
    L.drice_host_alloc.restype = vp
    L.drice_host_alloc.argtypes = [sz]
    L.drice_host_free.restype = None
    L.drice_host_free.argtypes = [vp]
    L.drice_encode_batch_dev_async.restype = i
    L.drice_encode_batch_dev_async.argtypes = [vp, vp, u64p, sz, i, i64, vp, sz, vp, vp, vp]
    L.drice_encode_batch_dev.restype = i
    L.drice_encode_batch_dev.argtypes = [vp, vp, u64p, sz, i, i64, vp, sz, u64p, vp]
    L.drice_decode_batch_dev_async.restype = i
    L.drice_decode_batch_dev_async.argtypes = [vp, vp, u64p, sz, u64p, i, i64, vp, vp, vp]
    L.drice_decode_batch_dev.restype = i
    L.drice_decode_batch_dev.argtypes = [vp, vp, u64p, sz, u64p, i, i64, vp, vp]
    L.drice_encode_batch_host.restype = i
    L.drice_encode_batch_host.argtypes = [vp, vp, u64p, sz, i, i64, vp, sz, u64p]
    L.drice_decode_batch_host.restype = i
    L.drice_decode_batch_host.argtypes = [vp, vp, u64p, sz, u64p, i, i64, vp]
    L.drice_peek_chunk_samples.restype = i
    L.drice_peek_chunk_samples.argtypes = [vp, u64p, sz, u64p]
    L.drice_launch_count.restype = C.c_uint64
    L.drice_launch_count.argtypes = [vp]
    L.drice_timing_enable.restype = i
    L.drice_timing_enable.argtypes = [vp, i]
    L.drice_timing_read.restype = i
    L.drice_timing_read.argtypes = [vp, C.POINTER(C.c_double), u64p, i]
    L.drice_kernel_name.restype = C.c_char_p
    L.drice_kernel_name.argtypes = [i]
    L.H5Z_filter_deltarice.restype = sz
    L.H5Z_filter_deltarice.argtypes = [C.c_uint, sz, C.POINTER(C.c_uint), sz, C.POINTER(sz), C.POINTER(vp)]
    L.deltarice_register_h5filter.restype = i
    L.H5PLget_plugin_type.restype = i
    L.H5PLget_plugin_info.restype = vp
    _lib = L
    return L
