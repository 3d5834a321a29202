"""CPU suite: the C-ABI library loads, exports every declared symbol and its host-side
logic (parameter parsing, bounds, error behaviour without a GPU) matches the reference's
parseCD_VALUES (reference src/deltaRice.c:248-291) and the format's size bound."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import deltarice_b200 as d
from deltarice_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "deltarice_b200.h")).read()
    declared = set(re.findall(r"\b(drice_\w+)\s*\(", hdr))
    declared -= {"drice_ctx"}
    assert declared == set(_lib.C_ABI_SYMBOLS)
    for s in list(declared) + _lib.H5_SYMBOLS:
        assert hasattr(L, s), s
    assert L.drice_abi_version() == 2


def test_headers_compile_as_c_and_cxx():
    """include/*.h is the drop-in boundary: plain C (what the reference's C callers compile) and C++."""
    import glob, shutil, subprocess
    if not shutil.which("gcc") or not shutil.which("g++"):
        pytest.skip("no host compiler")
    inc = os.path.join(ROOT, "include")
    shim = os.path.join(ROOT, "oracle", "hdf5_shim")       # (deltaRice.h includes <hdf5.h> under DRICE_USE_SYSTEM_HDF5 only)
    for h in sorted(glob.glob(os.path.join(inc, "*.h"))):
        for cc, std, lang in (("gcc", "-std=c99", "c"), ("g++", "-std=c++17", "c++")):
            r = subprocess.run([cc, std, "-Wall", "-Werror", "-fsyntax-only", "-x", lang, "-I", inc, "-I", shim, h],
                               capture_output=True, text=True)
            assert r.returncode == 0, f"{cc} {h}: {r.stderr[:500]}"


def test_h5_class_struct_matches_reference():
    # reference src/deltaRice.c:19-28: {vers 1, id 32025, enc 1, dec 1, "deltarice", NULL, NULL, filter}
    L = _lib.load()

    class H5ZClass2(C.Structure):
        _fields_ = [("version", C.c_int), ("id", C.c_int), ("enc", C.c_uint), ("dec", C.c_uint),
                    ("name", C.c_char_p), ("can_apply", C.c_void_p), ("set_local", C.c_void_p),
                    ("filter", C.c_void_p)]
    cls = H5ZClass2.in_dll(L, "H5Z_DELTARICE")
    assert (cls.version, cls.id, cls.enc, cls.dec, cls.name) == (1, 32025, 1, 1, b"deltarice")
    assert cls.can_apply is None and cls.set_local is None
    assert cls.filter == C.cast(L.H5Z_filter_deltarice, C.c_void_p).value
    # plugin entry points: type FILTER (0) and info -> the class (reference returns (void*)32025: bug B1)
    assert L.H5PLget_plugin_type() == 0
    assert L.H5PLget_plugin_info() == C.addressof(cls)


def test_parse_cd_values_defaults_and_forms():
    assert d.parse_cd_values(()) == (8, -1)
    assert d.parse_cd_values((16,)) == (16, -1)
    assert d.parse_cd_values((8, 1024)) == (8, 1024)
    assert d.parse_cd_values((4, 0xFFFFFFFF)) == (4, -1)
    assert d.parse_cd_values((8, 1024, 2, 1, 0xFFFFFFFF)) == (8, 1024)   # explicit delta filter


@pytest.mark.parametrize("cd", [(0,), (3,), (65536,), (8, 0), (8, 1024, 0), (8, 1024, 2, 1), (8, 1024, 2, 0, 1)])
def test_parse_cd_values_rejects(cd):
    with pytest.raises(d.DeltaRiceError):
        d.parse_cd_values(cd)


def test_parse_cd_values_generic_filter():
    """cd_nelmts >= 3 (reference src/deltaRice.c:277-290): ints arrive as two's-complement unsigned."""
    from deltarice_b200.codec import parse_cd_values_full
    assert parse_cd_values_full((8, 1024, 1, 1)) == (8, 1024, (1,))            # reference tests/test.py:48
    assert parse_cd_values_full((4, 0xFFFFFFFF, 3, 1, 0xFFFFFFFE, 1)) == (4, -1, (1, -2, 1))
    assert parse_cd_values_full((8,)) == (8, -1, (1, -1))
    with pytest.raises(d.DeltaRiceError) as e:                                 # more taps than the kernels take
        d.parse_cd_values((8, 1024, 17) + (1,) * 17)
    assert e.value.code == _lib.E_UNSUPPORTED


def test_log2_param():
    L = _lib.load()
    for k in range(16):
        assert L.drice_log2_param(1 << k) == k
    for M in (0, -4, 3, 12, 1 << 16, 1 << 20):
        assert L.drice_log2_param(M) == -1


def test_bound_matches_oracle(oracle):
    for total, L in [(0, None), (1, None), (32, None), (7000, 7000), (7001, 7000), (140000, 7000), (100, 7000), (65536, 1024), (50, 1)]:
        assert d.chunk_bound_bytes(total, L) == 4 * oracle.bound_words(total, L)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(d.DeltaRiceError) as e:
        d.DeltaRice(0)
    assert e.value.code == _lib.E_CUDA
    from deltarice_b200 import h5
    with pytest.raises(d.DeltaRiceError):
        h5.apply_filter(np.zeros(64, np.int16).tobytes(), (8, 32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "deltarice_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f


def test_chunk_grid_follows_hdf5_order():
    """deltarice_b200.h5.chunk_grid: chunk origins in libhdf5's row-major order, edge chunks included."""
    from deltarice_b200 import h5
    assert h5.chunk_grid((100, 7000), (20, 7000)) == [(0, 0), (20, 0), (40, 0), (60, 0), (80, 0)]
    g = h5.chunk_grid((50, 333), (16, 128))
    assert len(g) == 4 * 3 and g[0] == (0, 0) and g[1] == (0, 128) and g[3] == (16, 0) and g[-1] == (48, 256)
    assert h5.chunk_grid((7,), (3,)) == [(0,), (3,), (6,)]


def test_reference_module_name_imports():
    """Reference user scripts do `import deltaRice.h5` (README.md:65-91, tests/test.py:1-2): the shim
    package carries the reference module's names (src/h5.pyx:27, :55)."""
    import deltaRice.h5 as m
    assert m.H5FILTER == 32025
    assert callable(m.register_h5_filter)
    import deltarice_b200.h5 as h
    if not h.h5py_loaded():                              # no libhdf5 here: registration must FAIL loudly
        import pytest
        with pytest.raises(RuntimeError):
            m.register_h5_filter()
