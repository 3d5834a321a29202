"""the product behind a REAL libhdf5.  Runs wherever h5py is installed on a GPU box (this
image has neither h5py nor libhdf5: every test here skips); the first box that has them runs the
reference's own acceptance round trips (reference tests/test.py:8-83 — the same six option tuples and
data shapes, restated) through h5py in both ways the reference documents:

  * `import deltaRice.h5` registers the class with h5py's libhdf5 (reference README.md:65-91);
  * plugin discovery: HDF5_PLUGIN_PATH holds libh5deltarice_b200.so and H5PLget_plugin_info returns the
    class pointer (the reference returns the filter id, src/deltaRice_h5plugin.c:5) — checked in a
    subprocess that never imports the Python module.

The stored chunks must also be byte-equal to the oracle's stream for the same chunk."""
import os
import subprocess
import sys

import numpy as np
import pytest

h5py = pytest.importorskip("h5py", reason="no h5py / libhdf5 in this image (SURVEY 8c)")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [  # (name, data, compression_opts) — reference tests/test.py
    ("worst_case", lambda: np.random.default_rng(1).uniform(-32768, 32768, 2 ** 16).astype(np.int16), None),
    ("different_m", lambda: np.random.default_rng(2).uniform(-32768, 32768, 2 ** 16).astype(np.int16), (16,)),
    ("m_and_segment", lambda: np.random.default_rng(3).uniform(-32768, 32768, 2 ** 16).astype(np.int16), (8, 1024)),
    ("identity_filter", lambda: np.random.default_rng(4).uniform(-32768, 32768, 2 ** 16).astype(np.int16), (8, 1024, 1, 1)),
    ("all_signed", lambda: np.arange(-32768, 32768).astype(np.int16), (8, 1024, 1, 1)),
    ("all_unsigned", lambda: np.arange(0, 65536).astype(np.uint16), (8, 1024, 1, 1)),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_reference_round_trips_through_h5py(tmp_path, case):
    import deltaRice.h5                                   # registers filter 32025 at import
    from oracle import oracle as O
    name, make, opts = case
    data = make()
    path = str(tmp_path / f"{name}.h5")
    kw = {} if opts is None else {"compression_opts": opts}
    with h5py.File(path, "w") as f:
        f.create_dataset("test", data=data, compression=deltaRice.h5.H5FILTER, **kw)
    with h5py.File(path, "r") as f:
        d = f["test"]
        assert np.array_equal(d[()], data)
        mask, stored = d.id.read_direct_chunk((0,))
        assert mask == 0
    want = O.encode_chunk_cd(data.view(np.int16), tuple(opts or ()))
    assert np.array_equal(np.frombuffer(stored, np.uint32), want), "stored chunk differs from the oracle's stream"


def test_plugin_discovery_without_the_python_module(tmp_path):
    code = f"""
import h5py, numpy as np
x = np.random.default_rng(0).normal(0, 10, (100, 7000)).astype(np.int16)
with h5py.File(r'{tmp_path}/p.h5', 'w') as f:
    f.create_dataset('d', data=x, chunks=(20, 7000), compression=32025, compression_opts=(8, 7000))
with h5py.File(r'{tmp_path}/p.h5', 'r') as f:
    assert np.array_equal(f['d'][()], x)
print('ok')
"""
    env = dict(os.environ, HDF5_PLUGIN_PATH=os.path.join(ROOT, "deltarice_b200"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
