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
    """Returns H5Zregister's herr_t (< 0: failed / no libhdf5 loaded in this process).  The library looks
    for H5Zregister in the global scope first and then in every loaded object whose name contains
    "libhdf5" (h5py's extension modules pull libhdf5 in RTLD_LOCAL, wheels under a hashed name)."""
    return int(_lib.load().deltarice_register_h5filter())


def h5py_loaded() -> bool:
    """True when h5py can be imported (its libhdf5 is then in the process)."""
    try:
        import h5py  # noqa: F401
        return True
    except Exception:  # noqa: BLE001
        return False


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


# --------------------------------------------------------------------------------------
# direct-chunk batch front-end
# --------------------------------------------------------------------------------------
# libhdf5's filter pipeline hands the filter ONE chunk per call.  Writers / readers that own their
# chunking bypass it (H5Dwrite_chunk / H5Dread_chunk, h5py write_direct_chunk / read_direct_chunk)
# and push all the chunks of a dataset through the chunk scheduler in one call; the stored chunks
# are exactly what the stock CPU filter would have produced, so the file stays readable without
# this library.
def chunk_grid(shape, chunks):
    """Chunk origins of an HDF5 dataset of `shape` with chunk shape `chunks`, in the row-major
    order libhdf5 enumerates them; edge chunks are full size (HDF5 pads them)."""
    import itertools
    import math
    counts = [math.ceil(s / c) for s, c in zip(shape, chunks)]
    return [tuple(i * c for i, c in zip(idx, chunks)) for idx in itertools.product(*[range(n) for n in counts])]


def encode_dataset_chunks(codec, data, chunks, compression_opts=(), fill_value=0):
    """Encodes every chunk of an int16 array in ONE batch call.  Returns [(origin, bytes)] ready
    for `dset.id.write_direct_chunk(origin, bytes, filter_mask=0)`.  `compression_opts` is the
    reference's tuple (RiceParameter, WaveformLength[, filter_len, taps...]); WaveformLength
    counts samples of the flattened chunk, as in the filter."""
    import numpy as np
    from . import codec as _codec
    data = np.asarray(data)
    if data.dtype.itemsize != 2:
        raise ValueError("Delta-Rice codes 16-bit samples")
    M, L, taps = _codec.parse_cd_values_full(tuple(compression_opts))
    origins = chunk_grid(data.shape, chunks)
    n_per = int(np.prod(chunks))
    flat = np.empty(len(origins) * n_per, dtype=np.int16)
    for k, org in enumerate(origins):
        block = np.full(chunks, fill_value, dtype=data.dtype)
        src = tuple(slice(o, min(o + c, s)) for o, c, s in zip(org, chunks, data.shape))
        dst = tuple(slice(0, sl.stop - sl.start) for sl in src)
        block[dst] = data[src]
        flat[k * n_per:(k + 1) * n_per] = block.view(np.int16).ravel()
    off = np.arange(len(origins) + 1, dtype=np.uint64) * n_per
    codec.set_filter(None if taps == (1, -1) else taps)
    try:
        stream, boff = codec.encode_host(flat, off, M, None if L < 0 else L)
    finally:
        codec.set_filter(None)
    return [(org, stream[int(boff[k]):int(boff[k + 1])].tobytes()) for k, org in enumerate(origins)]


def decode_dataset_chunks(codec, chunk_bytes, shape, chunks, compression_opts=(), dtype="int16"):
    """Inverse of encode_dataset_chunks: `chunk_bytes` = the stored chunks in chunk_grid order
    (e.g. from `dset.id.read_direct_chunk(origin)[1]`); returns the array of `shape`."""
    import numpy as np
    from . import codec as _codec
    M, L, taps = _codec.parse_cd_values_full(tuple(compression_opts))
    origins = chunk_grid(shape, chunks)
    if len(chunk_bytes) != len(origins):
        raise ValueError("one stored chunk per grid cell expected")
    n_per = int(np.prod(chunks))
    comp = np.frombuffer(b"".join(chunk_bytes), dtype=np.uint8)
    boff = np.concatenate([[0], np.cumsum([len(b) for b in chunk_bytes])]).astype(np.uint64)
    off = np.arange(len(origins) + 1, dtype=np.uint64) * n_per
    codec.set_filter(None if taps == (1, -1) else taps)
    try:
        flat = codec.decode_host(comp, boff, off, M, None if L < 0 else L)
    finally:
        codec.set_filter(None)
    out = np.empty(shape, dtype=np.dtype(dtype))
    for k, org in enumerate(origins):
        block = flat[k * n_per:(k + 1) * n_per].view(out.dtype).reshape(chunks)
        dst = tuple(slice(o, min(o + c, s)) for o, c, s in zip(org, chunks, shape))
        src = tuple(slice(0, sl.stop - sl.start) for sl in dst)
        out[dst] = block[src]
    return out


def _auto_register():
    """With h5py in the process the filter is registered at import, as the reference module does
    (src/h5.pyx:61).  A failure is reported, not swallowed: HDF5_PLUGIN_PATH=plugin_path() is the route
    that does not depend on finding h5py's libhdf5."""
    if not h5py_loaded():
        return
    ret = -1
    try:
        ret = register_h5_filter()
    except Exception as e:  # noqa: BLE001
        import warnings
        warnings.warn(f"deltarice_b200.h5: registering filter {H5FILTER} failed ({e!r}); "
                      f"set HDF5_PLUGIN_PATH={plugin_path()} instead", RuntimeWarning)
        return
    if ret < 0:
        import warnings
        warnings.warn(f"deltarice_b200.h5: H5Zregister was not found or refused filter {H5FILTER} (ret {ret}); "
                      f"set HDF5_PLUGIN_PATH={plugin_path()} instead", RuntimeWarning)


_auto_register()
