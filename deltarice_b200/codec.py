"""codec.py — Python front-end of the C-ABI (include/deltarice_b200.h).

`DeltaRice` is a thin handle on a `drice_ctx`: it passes pointers to
libh5deltarice_b200.so and never touches sample data itself.  torch is used only for
device memory and streams ("plumbing"); numpy arrays go through the host-pointer entry
points (the chunk scheduler).

Argument meaning follows the reference's compression_opts tuple
(RiceParameter, WaveformLength) — reference README.md:69-91, src/deltaRice.c:248-291:
  M  RiceParameter, power of two          L  WaveformLength in samples, -1/None = whole chunk
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import DeltaRiceError

_U64P = C.POINTER(C.c_uint64)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def _p(a: np.ndarray):
    return a.ctypes.data_as(_U64P)


def _Lval(L) -> int:
    return -1 if (L is None or int(L) < 0) else int(L)


def chunk_offsets(chunk_samples: int | Sequence[int], total: int | None = None) -> np.ndarray:
    """Cumulative sample offsets [nchunks+1] from per-chunk sample counts, or from one chunk
    size and a total (last chunk short)."""
    if np.isscalar(chunk_samples):
        cs = int(chunk_samples)
        assert total is not None and cs > 0
        n = (int(total) + cs - 1) // cs
        off = np.minimum(np.arange(n + 1, dtype=np.uint64) * np.uint64(cs), np.uint64(total))
        return off
    cs = np.asarray(chunk_samples, dtype=np.uint64)
    return np.concatenate([np.zeros(1, np.uint64), np.cumsum(cs, dtype=np.uint64)])


def parse_cd_values_full(cd_values: Sequence[int] = ()) -> tuple[int, int, tuple]:
    """compression_opts -> (M, L, filter taps) exactly as the filter callback parses them
    (reference parseCD_VALUES, src/deltaRice.c:248-291)."""
    L = _lib.load()
    prm = _lib.Params()
    cd = (C.c_uint * max(1, len(cd_values)))(*[int(v) & 0xFFFFFFFF for v in cd_values])
    rc = L.drice_parse_cd_values(len(cd_values), cd, C.byref(prm))
    if rc != 0:
        raise DeltaRiceError(rc, f"bad compression_opts {tuple(cd_values)}")
    return int(prm.M), int(prm.L), tuple(int(prm.filter[i]) for i in range(prm.filter_len))


def parse_cd_values(cd_values: Sequence[int] = ()) -> tuple[int, int]:
    """compression_opts -> (M, L) exactly as the filter callback parses them."""
    L = _lib.load()
    prm = _lib.Params()
    cd = (C.c_uint * max(1, len(cd_values)))(*[int(v) & 0xFFFFFFFF for v in cd_values])
    rc = L.drice_parse_cd_values(len(cd_values), cd, C.byref(prm))
    if rc != 0:
        raise DeltaRiceError(rc, f"bad compression_opts {tuple(cd_values)}")
    return int(prm.M), int(prm.L)


def chunk_bound_bytes(nsamples: int, L=None) -> int:
    return int(_lib.load().drice_chunk_bound_bytes(int(nsamples), _Lval(L)))


class DeltaRice:
    """One codec context bound to one CUDA device (one per process/rank)."""

    def __init__(self, device: int | None = None):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.drice_create(C.byref(h), -1 if device is None else int(device))
        if rc != 0:
            raise DeltaRiceError(rc, (self._L.drice_last_error(None) or b"").decode())
        self._h = h
        self._pinned = []
        self.device = int(self._L.drice_device(h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.drice_destroy(self._h)
            self._h = None
            for p in self._pinned:
                self._L.drice_host_free(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise DeltaRiceError(rc, (self._L.drice_last_error(self._h) or b"").decode())

    def set_filter(self, taps=None):
        """Pre-filter of the following encode / decode calls (reference cd_values[3:]): None = the
        delta filter [1,-1] (fused in the kernels), (1,) = none, anything else = FIR before encode /
        recursion + division by taps[0] after decode (reference src/deltaRice.c:64-74, :91-102)."""
        if taps is None:
            self._check(self._L.drice_set_filter(self._h, None, 0))
        else:
            t = [int(v) for v in taps]
            arr = (C.c_int32 * max(1, len(t)))(*t)
            self._check(self._L.drice_set_filter(self._h, arr, len(t)))

    @property
    def launches(self) -> int:
        return int(self._L.drice_launch_count(self._h))

    def timing(self, on: bool = True):
        """Bracket every kernel launch with CUDA events (bench.py's roofline leg)."""
        self._check(self._L.drice_timing_enable(self._h, 1 if on else 0))

    def timing_read(self, reset: bool = True) -> dict:
        """{kernel name: (summed ms, launches)} since the last reset; waits for pending events."""
        n = 3
        ms = (C.c_double * n)()
        cnt = (C.c_uint64 * n)()
        self._check(self._L.drice_timing_read(self._h, ms, cnt, 1 if reset else 0))
        return {self._L.drice_kernel_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n)}

    def pinned_empty(self, n: int, dtype) -> np.ndarray:
        """numpy array over page-locked host memory (drice_host_alloc); freed with the codec."""
        dt = np.dtype(dtype)
        nbytes = max(1, int(n) * dt.itemsize)
        p = self._L.drice_host_alloc(nbytes)
        if not p:
            raise DeltaRiceError(_lib.E_NOMEM, "drice_host_alloc failed")
        self._pinned.append(p)
        buf = (C.c_uint8 * nbytes).from_address(p)
        return np.frombuffer(buf, dtype=dt, count=int(n))

    def bound_bytes(self, sample_off, L=None) -> int:
        off = _u64(sample_off)
        return int(self._L.drice_batch_bound_bytes(_p(off), off.size - 1, _Lval(L)))

    # ------------------------------------------------------------------ host buffers
    def encode_host(self, raw: np.ndarray, sample_off=None, M: int = 8, L=None):
        """int16 host array -> (uint8 stream of all chunks, byte offsets [nchunks+1])."""
        raw = np.ascontiguousarray(raw)
        if raw.dtype.itemsize != 2:
            raise DeltaRiceError(_lib.E_PARAM, "samples must be 16-bit")
        raw = raw.reshape(-1)
        off = _u64([0, raw.size] if sample_off is None else sample_off)
        n = off.size - 1
        cap = self.bound_bytes(off, L)
        out = np.empty(cap, dtype=np.uint8)
        boff = np.zeros(n + 1, dtype=np.uint64)
        self._check(self._L.drice_encode_batch_host(self._h, raw.ctypes.data, _p(off), n, int(M), _Lval(L),
                                                    out.ctypes.data, cap, _p(boff)))
        return out[: int(boff[-1])], boff

    def encode_host_into(self, raw: np.ndarray, sample_off, M: int, L, out: np.ndarray, boff: np.ndarray) -> int:
        """As encode_host, into caller buffers (e.g. pinned_empty); returns the stream's bytes."""
        off = _u64(sample_off)
        self._check(self._L.drice_encode_batch_host(self._h, raw.ctypes.data, _p(off), off.size - 1, int(M), _Lval(L),
                                                    out.ctypes.data, out.nbytes, _p(boff)))
        return int(boff[off.size - 1])

    def decode_host_into(self, comp: np.ndarray, byte_off, sample_off, M: int, L, out: np.ndarray) -> None:
        boff, off = _u64(byte_off), _u64(sample_off)
        self._check(self._L.drice_decode_batch_host(self._h, comp.ctypes.data, _p(boff), boff.size - 1, _p(off),
                                                    int(M), _Lval(L), out.ctypes.data))

    def decode_host(self, comp: np.ndarray, byte_off=None, sample_off=None, M: int = 8, L=None) -> np.ndarray:
        comp = np.ascontiguousarray(comp).view(np.uint8).reshape(-1)
        boff = _u64([0, comp.size] if byte_off is None else byte_off)
        n = boff.size - 1
        if sample_off is None:
            cs = np.zeros(n, dtype=np.uint64)
            rc = self._L.drice_peek_chunk_samples(comp.ctypes.data, _p(boff), n, _p(cs))
            if rc != 0:
                raise DeltaRiceError(rc, "chunk shorter than its header")
            off = chunk_offsets(cs)
        else:
            off = _u64(sample_off)
        out = np.empty(int(off[-1]), dtype=np.int16)
        self._check(self._L.drice_decode_batch_host(self._h, comp.ctypes.data, _p(boff), n, _p(off), int(M),
                                                    _Lval(L), out.ctypes.data))
        return out

    # ------------------------------------------------------------------ device buffers (torch)
    @staticmethod
    def _stream():
        import torch
        h = torch.cuda.current_stream().cuda_stream
        # torch's default stream is the legacy NULL stream; NULL means "the context's own
        # stream" in the C-ABI, so name the legacy stream explicitly (cudaStreamLegacy == 1)
        return C.c_void_p(h if h else 1)

    def encode_device(self, raw, sample_off=None, M: int = 8, L=None, out=None):
        """torch int16 CUDA tensor -> (torch uint8 CUDA tensor [valid bytes], byte offsets ndarray)."""
        import torch
        assert raw.is_cuda and raw.element_size() == 2 and raw.is_contiguous()
        off = _u64([0, raw.numel()] if sample_off is None else sample_off)
        n = off.size - 1
        cap = self.bound_bytes(off, L)
        if out is None:
            out = torch.empty(cap, dtype=torch.uint8, device=raw.device)
        boff = np.zeros(n + 1, dtype=np.uint64)
        self._check(self._L.drice_encode_batch_dev(self._h, raw.data_ptr(), _p(off), n, int(M), _Lval(L),
                                                   out.data_ptr(), out.numel() * out.element_size(), _p(boff),
                                                   self._stream()))
        return out.view(torch.uint8)[: int(boff[-1])], boff

    def encode_device_async(self, raw, sample_off, M, L, out, d_byte_off, d_status):
        """Enqueue only (no sync): offsets and status stay in device tensors."""
        off = _u64(sample_off)
        self._check(self._L.drice_encode_batch_dev_async(self._h, raw.data_ptr(), _p(off), off.size - 1, int(M),
                                                         _Lval(L), out.data_ptr(), out.numel() * out.element_size(),
                                                         d_byte_off.data_ptr(), d_status.data_ptr(), self._stream()))

    def decode_device(self, comp, byte_off, sample_off, M: int = 8, L=None, out=None):
        import torch
        assert comp.is_cuda and comp.is_contiguous()
        boff, off = _u64(byte_off), _u64(sample_off)
        if out is None:
            out = torch.empty(int(off[-1]), dtype=torch.int16, device=comp.device)
        self._check(self._L.drice_decode_batch_dev(self._h, comp.data_ptr(), _p(boff), boff.size - 1, _p(off),
                                                   int(M), _Lval(L), out.data_ptr(), self._stream()))
        return out

    def decode_device_async(self, comp, byte_off, sample_off, M, L, out, d_status):
        boff, off = _u64(byte_off), _u64(sample_off)
        self._check(self._L.drice_decode_batch_dev_async(self._h, comp.data_ptr(), _p(boff), boff.size - 1, _p(off),
                                                         int(M), _Lval(L), out.data_ptr(), d_status.data_ptr(),
                                                         self._stream()))
