"""oracle.py — Python face of the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs)
may import this module.  Nothing in deltarice_b200/ does.

Three independent checkers of the Delta-Rice stream (reference src/deltaRice.c):
  * `encode_chunk` / `decode_chunk`   — C restatement, oracle/drice_oracle.c (libdrice_oracle.so)
  * `np_encode_chunk` / `np_decode_chunk` — numpy/pure-Python restatement (small inputs)
  * `ref_filter`                       — the UNMODIFIED reference compiled into oracle/_ref/
                                         (present where `make -C oracle ref` has run; the
                                         prebuilt .so travels to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libdrice_oracle.so")
_REF_DIR = os.path.join(_HERE, "_ref")
H5Z_FLAG_REVERSE = 0x0100


def build(ref: bool | None = None) -> None:
    """Compile the restatement and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    if ref is None:
        ref = os.path.exists("/root/reference/src/deltaRice.c")
    if ref:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build(ref=False)
        L = C.CDLL(_LIB)
        sz, vp, i = C.c_size_t, C.c_void_p, C.c_int
        L.drice_oracle_log2_param.restype = i
        L.drice_oracle_log2_param.argtypes = [i]
        L.drice_oracle_chunk_bound_words.restype = sz
        L.drice_oracle_chunk_bound_words.argtypes = [sz, sz]
        for name in ("drice_oracle_encode_chunk", "drice_oracle_encode_chunk_mt",
                     "drice_oracle_decode_chunk", "drice_oracle_decode_chunk_mt"):
            f = getattr(L, name)
            f.restype = sz
            f.argtypes = [vp, sz, i, sz, vp, sz]
        for name in ("drice_oracle_encode_chunk_f", "drice_oracle_decode_chunk_f"):
            f = getattr(L, name)
            f.restype = sz
            f.argtypes = [vp, sz, i, sz, vp, i, vp, sz]
        _lib = L
    return _lib


def _L(L: int | None) -> int:
    return 0 if (L is None or L < 0) else int(L)


def bound_words(total: int, L: int | None) -> int:
    return int(lib().drice_oracle_chunk_bound_words(total, _L(L)))


def _filt(filt):
    f = np.ascontiguousarray(np.asarray(filt, dtype=np.int64).astype(np.int32))
    return f, f.ctypes.data, int(f.size)


def cd_values(M: int, L: int | None, filt=None) -> tuple:
    """compression_opts tuple as h5py would pass it (ints as two's-complement unsigned;
    reference parseCD_VALUES src/deltaRice.c:248-291): (M, L[, filter_len, f0, f1, ...])."""
    if filt is None:
        return (M,) if _L(L) == 0 else (M, _L(L))
    Lv = 0xFFFFFFFF if _L(L) == 0 else _L(L)
    return (M, Lv, len(filt)) + tuple(int(v) & 0xFFFFFFFF for v in filt)


def encode_chunk(x: np.ndarray, M: int = 8, L: int | None = None, mt: bool = False, filt=None) -> np.ndarray:
    """int16[total] -> uint32 stream words of one chunk (reference src/deltaRice.c:383-436).
    `filt`: pre-filter taps (cd_values[3:], reference :64-74); None = the delta filter [1,-1]."""
    x = np.ascontiguousarray(x).view(np.int16).ravel()
    cap = bound_words(x.size, L)
    out = np.empty(cap, dtype=np.uint32)
    if filt is not None:
        f, fp, fl = _filt(filt)
        n = lib().drice_oracle_encode_chunk_f(x.ctypes.data, x.size, int(M), _L(L), fp, fl, out.ctypes.data, cap)
        if n == 0:
            raise ValueError(f"oracle encode rejected M={M} L={L} total={x.size} filt={filt}")
        return out[:n].copy()
    fn = lib().drice_oracle_encode_chunk_mt if mt else lib().drice_oracle_encode_chunk
    n = fn(x.ctypes.data, x.size, int(M), _L(L), out.ctypes.data, cap)
    if n == 0:
        raise ValueError(f"oracle encode rejected M={M} L={L} total={x.size}")
    return out[:n].copy()


def encode_chunk_cd(x: np.ndarray, cd=()) -> np.ndarray:
    """encode_chunk driven by a compression_opts tuple as h5py passes it (parseCD_VALUES,
    src/deltaRice.c:248-291): () -> M=8, whole chunk; (M,); (M, L); (M, L, filter_len, taps...)."""
    cd = tuple(int(v) for v in cd)
    M = cd[0] if len(cd) >= 1 else 8
    L = None
    if len(cd) >= 2:
        Lv = cd[1] & 0xFFFFFFFF
        L = None if Lv == 0xFFFFFFFF else Lv
    filt = None
    if len(cd) >= 3:
        filt = [((v & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000 for v in cd[3:3 + cd[2]]]
        if filt == [1, -1]:
            filt = None
    return encode_chunk(x, M, L, filt=filt)


def decode_chunk(words: np.ndarray, M: int = 8, L: int | None = None, mt: bool = False, filt=None) -> np.ndarray:
    """uint32 stream words of one chunk -> int16[total] (reference src/deltaRice.c:301-341).
    `filt`: pre-filter taps to invert (reference :91-102); None = the delta filter."""
    w = np.ascontiguousarray(words).view(np.uint32).ravel()
    if w.size < 1:
        raise ValueError("empty stream")
    total = int(w[0])
    y = np.empty(total, dtype=np.int16)
    if filt is not None:
        f, fp, fl = _filt(filt)
        n = lib().drice_oracle_decode_chunk_f(w.ctypes.data, w.size, int(M), _L(L), fp, fl, y.ctypes.data, total)
    else:
        fn = lib().drice_oracle_decode_chunk_mt if mt else lib().drice_oracle_decode_chunk
        n = fn(w.ctypes.data, w.size, int(M), _L(L), y.ctypes.data, total)
    if n == C.c_size_t(-1).value:
        raise ValueError("oracle decode: malformed stream")
    return y


# --------------------------------------------------------------------------------------
# numpy / pure-Python restatement (independent of the C one; small inputs only)
# --------------------------------------------------------------------------------------
def np_code_table(x: np.ndarray, k: int):
    """Per-sample (value, length) of one wave: delta (:53-62), zig-zag (:207-211),
    Rice / escape split (:212-228)."""
    x = x.astype(np.int16)
    d = np.empty_like(x)
    if x.size:
        d[0] = x[0]
        d[1:] = (x[1:].astype(np.int32) - x[:-1].astype(np.int32)).astype(np.int16)
    d32 = d.astype(np.int64)
    u = np.where(d32 >= 0, 2 * d32, -2 * d32 - 1).astype(np.int64)
    q = u >> k
    esc = q >= 8
    val = np.where(esc, (1 << 16) | u, (1 << k) | (u & ((1 << k) - 1)))
    ln = np.where(esc, 25, q + 1 + k)
    return val, ln


def np_encode_wave(x: np.ndarray, k: int) -> np.ndarray:
    val, ln = np_code_table(x, k)
    bits = []
    for v, n in zip(val.tolist(), ln.tolist()):
        bits.extend((v >> (n - 1 - b)) & 1 for b in range(n))
    pad = (-len(bits)) % 32
    bits.extend([0] * pad)
    if not bits:
        return np.zeros(0, dtype=np.uint32)
    by = np.packbits(np.array(bits, dtype=np.uint8))          # MSB-first bytes
    return by.reshape(-1, 4).view(">u4").astype(np.uint32).ravel()


def np_encode_chunk(x: np.ndarray, M: int = 8, L: int | None = None) -> np.ndarray:
    x = np.ascontiguousarray(x).view(np.int16).ravel()
    k = int(M).bit_length() - 1
    assert M > 0 and (1 << k) == M and k <= 15
    total = x.size
    Lw = total if _L(L) == 0 else _L(L)
    out = [np.array([total], dtype=np.uint32)]
    s = 0
    while s < total:
        w = np_encode_wave(x[s:s + Lw], k)
        out.append(np.array([w.size], dtype=np.uint32))
        out.append(w)
        s += Lw
    return np.concatenate(out)


def np_decode_chunk(words: np.ndarray, M: int = 8, L: int | None = None) -> np.ndarray:
    w = np.ascontiguousarray(words).view(np.uint32).ravel()
    k = int(M).bit_length() - 1
    total = int(w[0])
    Lw = total if _L(L) == 0 else _L(L)
    y = np.empty(total, dtype=np.int16)
    cur, s = 1, 0
    while s < total:
        n = min(Lw, total - s)
        nw = int(w[cur])
        bits = np.unpackbits(w[cur + 1:cur + 1 + nw].astype(">u4").view(np.uint8))
        p, acc = 0, 0
        for i in range(n):
            q = 0
            while bits[p] == 0:
                q += 1
                p += 1
            p += 1
            nb = 16 if q == 8 else k
            v = 0
            for _ in range(nb):
                v = (v << 1) | int(bits[p])
                p += 1
            u = v if q == 8 else (q << k) + v
            d = -((u + 1) >> 1) if (u & 1) else (u >> 1)
            acc = d if i == 0 else acc + d
            acc = ((acc + 32768) & 0xFFFF) - 32768
            y[s + i] = acc
        cur += nw + 1
        s += n
    return y


# --------------------------------------------------------------------------------------
# the unmodified reference (oracle/_ref), driven through its own H5Z entry point
# --------------------------------------------------------------------------------------
_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]
_libc.free.argtypes = [C.c_void_p]
_ref_libs: dict[str, C.CDLL] = {}


def _cpu_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return " avx2" in f.read()
    except OSError:
        return False


def ref_path(kind: str = "omp") -> str | None:
    """kind: 'omp' (semantic oracle, handles the short last wave) or 'ser' (serial build;
    only valid when WaveformLength divides the chunk)."""
    names = {"omp": ["libref_omp.so"] if _cpu_has_avx2() else [], "ser": ["libref_ser.so"] if _cpu_has_avx2() else []}
    names["omp"].append("libref_omp_generic.so")
    for n in names[kind]:
        p = os.path.join(_REF_DIR, n)
        if os.path.exists(p):
            return p
    return None


def ref_available(kind: str = "omp") -> bool:
    return ref_path(kind) is not None


def ref_lib(kind: str = "omp") -> C.CDLL:
    if kind not in _ref_libs:
        p = ref_path(kind)
        if p is None:
            raise FileNotFoundError("oracle/_ref is not built (make -C oracle ref needs /root/reference)")
        L = C.CDLL(p)
        L.H5Z_filter_deltarice.restype = C.c_size_t
        L.H5Z_filter_deltarice.argtypes = [C.c_uint, C.c_size_t, C.POINTER(C.c_uint), C.c_size_t,
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
        _ref_libs[kind] = L
    return _ref_libs[kind]


def ref_filter(data: bytes | np.ndarray, cd_values=(), reverse: bool = False, kind: str = "omp") -> bytes:
    """One call of the reference's H5Z_filter_deltarice (src/deltaRice.c:468-490) with
    libc-malloc'ed buffers, exactly as libhdf5's pipeline would drive it."""
    raw = data.tobytes() if isinstance(data, np.ndarray) else bytes(data)
    L = ref_lib(kind)
    n = len(raw)
    p = _libc.malloc(max(n, 8) + 64)          # decoder reads one word past a record (B8)
    C.memmove(p, raw, n)
    C.memset(p + n, 0, 64)
    buf = C.c_void_p(p)
    buf_size = C.c_size_t(n)
    cd = (C.c_uint * max(1, len(cd_values)))(*[v & 0xFFFFFFFF for v in cd_values])
    ret = L.H5Z_filter_deltarice(H5Z_FLAG_REVERSE if reverse else 0, len(cd_values), cd, n,
                                 C.byref(buf_size), C.byref(buf))
    if ret in (0, C.c_size_t(-1).value):
        raise RuntimeError("reference filter failed")
    out = C.string_at(buf.value, ret)
    _libc.free(buf)
    return out


def ref_encode_chunk(x: np.ndarray, M: int = 8, L: int | None = None, kind: str = "omp", filt=None) -> np.ndarray:
    x = np.ascontiguousarray(x).view(np.int16).ravel()
    cd = cd_values(M, L, filt)
    return np.frombuffer(ref_filter(x, cd, False, kind), dtype=np.uint32).copy()


def ref_decode_chunk(words: np.ndarray, M: int = 8, L: int | None = None, kind: str = "omp", filt=None) -> np.ndarray:
    w = np.ascontiguousarray(words).view(np.uint32).ravel()
    cd = cd_values(M, L, filt)
    return np.frombuffer(ref_filter(w, cd, True, kind), dtype=np.int16).copy()
