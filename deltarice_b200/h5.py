"""h5.py — Python mirror of the reference's `deltaRice.h5` module (reference src/h5.pyx).

  H5FILTER              filter id 32025                        (src/h5.pyx:27, src/deltaRice.h:7)
  register_h5_filter()  H5Zregister(H5Z_DELTARICE) in-process   (src/h5.pyx:55-61)
  plugin_path()         directory to put on HDF5_PLUGIN_PATH    (setup.py --h5plugin-dir)
  apply_filter()        drives H5Z_filter_deltarice exactly as libhdf5's pipeline does
                        (malloc'ed *buf, ownership hand-over); used where no libhdf5 /
                        h5py is installed (this image) and by the parity tests.

With h5py installed, `import deltarice_b200.h5` registers the filter at import like the
reference module does, so `f.create_dataset(..., compression=H5FILTER,
compression_opts=(M, L))` (reference README.md:69-91) works unchanged.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _lib

H5FILTER = 32025
H5Z_FLAG_REVERSE = 0x0100

_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]
_libc.free.argtypes = [C.c_void_p]


def plugin_path() -> str:
    return os.path.dirname(_lib.LIB_PATH)


def register_h5_filter() -> int:
    """Returns H5Zregister's herr_t (< 0: failed / no libhdf5 loaded in this process)."""
    return int(_lib.load().deltarice_register_h5filter())


def apply_filter(data: bytes, cd_values=(), reverse: bool = False) -> bytes:
    """One H5Z_filter_deltarice call with libc-malloc'ed buffers. Raises on filter failure (0)."""
    L = _lib.load()
    raw = bytes(data)
    n = len(raw)
    p = _libc.malloc(max(n, 1))
    C.memmove(p, raw, n)
    buf = C.c_void_p(p)
    buf_size = C.c_size_t(n)
    cd = (C.c_uint * max(1, len(cd_values)))(*[int(v) & 0xFFFFFFFF for v in cd_values])
    ret = L.H5Z_filter_deltarice(H5Z_FLAG_REVERSE if reverse else 0, len(cd_values), cd, n,
                                 C.byref(buf_size), C.byref(buf))
    if ret == 0:
        _libc.free(buf)          # untouched input: still ours
        raise _lib.DeltaRiceError(_lib.E_PARAM, "H5Z_filter_deltarice returned 0 (failure)")
    out = C.string_at(buf.value, ret)
    assert buf_size.value == ret
    _libc.free(buf)
    return out


def _auto_register():
    try:
        import h5py  # noqa: F401
    except Exception:
        return
    try:
        register_h5_filter()
    except Exception:
        pass


_auto_register()
