"""GPU parity suite (-m gpu): the CUDA path, called through the C-ABI (ctypes ->
libh5deltarice_b200.so), against the oracle on the same seeded inputs.  Bit-exact bar:
the compressed stream must equal the oracle's byte for byte, decode must reproduce the
input exactly.  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

from cases import small_cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz")


def _oracle_batch(oracle, x, off, M, L):
    parts, boff = [], [0]
    for c in range(len(off) - 1):
        s = oracle.encode_chunk(x[int(off[c]):int(off[c + 1])], M, L)
        parts.append(s)
        boff.append(boff[-1] + 4 * s.size)
    return (np.concatenate(parts) if parts else np.zeros(0, np.uint32)), np.array(boff, dtype=np.uint64)


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_single_chunk_host_path(codec, oracle, case):
    name, x, M, L = case
    x = x.view(np.int16)
    want = oracle.encode_chunk(x, M, L)
    got, boff = codec.encode_host(x, None, M, L)
    assert int(boff[-1]) == 4 * want.size, name
    assert np.array_equal(got.view(np.uint32), want), name
    back = codec.decode_host(want.view(np.uint8), None, None, M, L)
    assert np.array_equal(back, x), name


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_single_chunk_h5z_filter(oracle, case):
    """Through H5Z_filter_deltarice with malloc'ed buffers, as libhdf5 drives it
    (reference src/deltaRice.c:468-490)."""
    from deltarice_b200 import h5
    name, x, M, L = case
    x = x.view(np.int16)
    cd = (M,) if L is None else (M, L)
    stream = h5.apply_filter(x.tobytes(), cd, reverse=False)
    assert np.array_equal(np.frombuffer(stream, np.uint32), oracle.encode_chunk(x, M, L)), name
    back = h5.apply_filter(stream, cd, reverse=True)
    assert np.array_equal(np.frombuffer(back, np.int16), x), name


def test_golden_vectors_through_filter():
    """Streams produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
    from deltarice_b200 import h5
    g = np.load(GOLDEN)
    for name in g["names"]:
        x, cd, stream = g[f"{name}__x"], tuple(int(v) for v in g[f"{name}__cd"]), g[f"{name}__stream"]
        got = h5.apply_filter(x.tobytes(), cd, reverse=False)
        assert np.array_equal(np.frombuffer(got, np.uint32), stream), name
        back = h5.apply_filter(stream.tobytes(), cd, reverse=True)
        assert np.array_equal(np.frombuffer(back, np.int16), x), name


def _cd_split(cd):
    cd = [int(v) for v in cd]
    M = cd[0] if len(cd) >= 1 else 8
    L = cd[1] if len(cd) >= 2 else None
    if L is not None and L >= 0x80000000:
        L = None
    taps = None
    if len(cd) >= 3:
        taps = [v - (1 << 32) if v >= 0x80000000 else v for v in cd[3:3 + cd[2]]]
    return M, L, taps


def test_golden_filter_vectors_through_filter():
    """compression_opts with a pre-filter (cd_nelmts >= 3, incl. the reference's own
    tests/test.py:46-83 option tuple (8, 1024, 1, 1)): streams and decoded outputs of the
    UNMODIFIED reference (tests/golden/make_golden_filters.py), through H5Z_filter_deltarice."""
    from deltarice_b200 import h5
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "golden_filters_v1.npz"))
    for name in g["names"]:
        x, cd = g[f"{name}__x"], tuple(int(v) for v in g[f"{name}__cd"])
        stream, back = g[f"{name}__stream"], g[f"{name}__back"]
        got = h5.apply_filter(x.tobytes(), cd, reverse=False)
        assert np.array_equal(np.frombuffer(got, np.uint32), stream), name
        dec = h5.apply_filter(stream.tobytes(), cd, reverse=True)
        assert np.array_equal(np.frombuffer(dec, np.int16), back), name
    # and the default filter still works on the shared handle afterwards
    x = g["second_difference__x"]
    s = h5.apply_filter(x.tobytes(), (4, 7000), reverse=False)
    assert h5.apply_filter(s, (4, 7000), reverse=True) == x.tobytes()


@pytest.mark.parametrize("taps", [[1], [1, -2, 1], [-1, 1], [2, -1], [3, 1, -2, 5], [1, 0, 0, 0, 0, 0, 0, -1]],
                         ids=lambda t: "f" + "_".join(str(v) for v in t))
@pytest.mark.parametrize("M,L", [(8, 7000), (4, 3500), (2, 33), (16, None), (64, 8178)])
def test_generic_filter_batch(codec, oracle, taps, M, L):
    """Ragged multi-chunk batches with a pre-filter against the oracle (FIR before encode,
    recursion + division after decode; reference src/deltaRice.c:64-74, :91-102)."""
    rng = np.random.default_rng(len(taps) * 100 + M)
    sizes = [7000 * 3, 0, 1, 3500 * 2 + 17, 8178 * 2, 64]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    x = np.cumsum(rng.normal(0, 9, int(off[-1]))).astype(np.int16)
    x[:5000] = rng.integers(-32768, 32768, 5000)                  # escapes and wrap-around
    codec.set_filter(taps)
    try:
        got, boff = codec.encode_host(x, off, M, L)
        parts = [oracle.encode_chunk(x[int(off[c]):int(off[c + 1])], M, L, filt=taps) for c in range(len(sizes))]
        want = np.concatenate(parts)
        assert int(boff[-1]) == 4 * want.size
        assert np.array_equal(got.view(np.uint32), want)
        back = codec.decode_host(got, boff, off, M, L)
        want_back = np.concatenate([oracle.decode_chunk(p, M, L, filt=taps) for p in parts])
        assert np.array_equal(back, want_back)
        if abs(taps[0]) == 1:
            assert np.array_equal(back, x)
    finally:
        codec.set_filter(None)
    # the default delta filter is back
    got, boff = codec.encode_host(x[:7000], None, M, L)
    assert np.array_equal(got.view(np.uint32), oracle.encode_chunk(x[:7000], M, L))


def test_set_filter_rejects(codec):
    import deltarice_b200 as d
    for bad in ([], [0, 1], [1] * 17):
        with pytest.raises(d.DeltaRiceError):
            codec.set_filter(bad)
    codec.set_filter(None)


def test_arithmetic_front_end_on_the_small_cases():
    """For RiceParameter 2, 4 and 8 the tile encoder looks pair codes up in a shared-memory table; every
    other parameter (and pre-filtered input) takes the arithmetic front-end.  DRICE_ENC_LUT=0 (read once per
    process) sends EVERY batch through the arithmetic one: the edge-case suite must stay bit-exact there too."""
    import subprocess
    import sys
    if os.environ.get("DRICE_ENC_LUT"):
        pytest.skip("already running under the override")
    env = dict(os.environ, DRICE_ENC_LUT="0")
    sel = ("single_chunk_host_path or single_chunk_h5z_filter or golden or ragged_batch or generic_filter_batch "
           "or unaligned_pointers or many_chunks or readme_config or errors or full_size_c2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k", sel],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("workers,stage", [(12, 0), (24, 0), (8, 0), (12, 300), (8, 64)])
def test_tile_geometries_and_small_staging(workers, stage):
    """The tile encoder runs as 2 x 12, 1 x 24 or 2 x 8 worker warps per SM depending on the staging a wave
    needs, and a wave that outgrows its staging is packed a second time straight into its record: force
    each geometry (DRICE_ENC_WORKERS) and a staging far too small (DRICE_ENC_STAGE_WORDS) on cases of
    every Rice parameter class."""
    import subprocess
    import sys
    if os.environ.get("DRICE_ENC_WORKERS"):
        pytest.skip("already running under the override")
    env = dict(os.environ, DRICE_ENC_WORKERS=str(workers))
    if stage:
        env["DRICE_ENC_STAGE_WORDS"] = str(stage)
    sel = "golden or ragged_batch or unaligned_pointers or readme_config or rice_parameter_sweep"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k", sel],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout


def test_header_chase_on_the_small_cases():
    """Records are located by scanning for headers when the waves are long enough, by chasing the
    chain (locate_kernel) otherwise.  DRICE_LOCATE_SCAN=0 sends every batch to the chase."""
    import subprocess
    import sys
    if os.environ.get("DRICE_LOCATE_SCAN"):
        pytest.skip("already running under the override")
    env = dict(os.environ, DRICE_LOCATE_SCAN="0")
    sel = ("single_chunk_host_path or golden or ragged_batch or unaligned_pointers or many_chunks or readme_config "
           "or errors or full_size_c2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k", sel],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout


def test_direct_chase_on_the_small_cases():
    """Very large batches locate their records with one warp per chunk walking the chain through global memory
    (chase_direct_kernel) instead of the header scan.  DRICE_LOCATE_DIRECT=1 sends every scannable batch there."""
    import subprocess
    import sys
    if os.environ.get("DRICE_LOCATE_DIRECT"):
        pytest.skip("already running under the override")
    env = dict(os.environ, DRICE_LOCATE_DIRECT="1")
    sel = ("single_chunk_host_path or golden or ragged_batch or unaligned_pointers or many_chunks or readme_config "
           "or errors or full_size_c2 or mixed_density")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k", sel],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout


def test_lane_parser_on_the_small_cases():
    """Batches of up to 896 waves (<= 8192 samples each) are decoded by parse_wide_kernel (one CTA
    per wave, parallel inside the wave), larger ones by parse_kernel (one lane per wave).
    DRICE_PARSE_WIDE=0 (read once per process) sends EVERY batch to the lane kernel: the edge-case
    suite must stay exact there too."""
    import subprocess
    import sys
    if os.environ.get("DRICE_PARSE_WIDE"):
        pytest.skip("already running under the override")
    env = dict(os.environ, DRICE_PARSE_WIDE="0")
    sel = ("single_chunk_host_path or single_chunk_h5z_filter or golden or ragged_batch or generic_filter_batch "
           "or unaligned_pointers or many_chunks or readme_config or errors")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k", sel],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout


def test_readme_config_c1_batch(codec, oracle):
    """C1: (100,7000) N(0,10), M=8, chunks (20,7000): 5 chunks in ONE launch."""
    x = np.random.default_rng(0).normal(0, 10, (100, 7000)).astype(np.int16).ravel()
    off = np.arange(6, dtype=np.uint64) * 140000
    want, wboff = _oracle_batch(oracle, x, off, 8, 7000)
    got, boff = codec.encode_host(x, off, 8, 7000)
    assert np.array_equal(boff, wboff)
    assert np.array_equal(got.view(np.uint32), want)
    ratio = got.size / (x.size * 2)
    assert abs(ratio - 0.405) < 0.002          # 6.48 bits/sample
    back = codec.decode_host(got, boff, off, 8, 7000)
    assert np.array_equal(back, x)


@pytest.mark.parametrize("M,L", [(8, 7000), (4, 3500), (2, 33), (16, None), (1, 64)])
def test_ragged_batch_with_empty_chunk(codec, oracle, M, L):
    r = np.random.default_rng(11)
    sizes = [7000 * 3, 0, 12345, 1, 7000 * 2 + 1, 3500, 0, 64]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    sig = 1.0 if M == 1 else 30.0
    x = np.clip(np.rint(r.normal(0, sig, int(off[-1]))), -32768, 32767).astype(np.int16)
    want, wboff = _oracle_batch(oracle, x, off, M, L)
    got, boff = codec.encode_host(x, off, M, L)
    assert np.array_equal(boff, wboff)
    assert np.array_equal(got.view(np.uint32), want)
    assert np.array_equal(codec.decode_host(got, boff, off, M, L), x)
    assert np.array_equal(codec.decode_host(got, boff, None, M, L), x)    # sizes peeked from the stream


def test_device_path_unaligned_pointers(codec, oracle):
    import torch
    r = np.random.default_rng(21)
    for shift in (0, 1, 3, 5, 8):
        for (M, L, n) in [(4, 3500, 3500 * 9), (8, 7000, 7000 * 4 + 17), (8, 1023, 1023 * 11), (2, None, 30000)]:
            x = np.clip(np.rint(r.normal(0, 25, n)), -32768, 32767).astype(np.int16)
            buf = torch.zeros(n + 64, dtype=torch.int16, device="cuda")
            d = buf[shift:shift + n]
            d.copy_(torch.from_numpy(x))
            off = np.array([0, n], dtype=np.uint64)
            cap = codec.bound_bytes(off, L)
            obuf = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
            o = obuf[4 * (shift % 4):4 * (shift % 4) + cap]
            comp, boff = codec.encode_device(d, off, M, L, out=o)
            want = oracle.encode_chunk(x, M, L)
            assert np.array_equal(comp.cpu().numpy().view(np.uint32), want), (shift, M, L)
            dec_buf = torch.zeros(n + 64, dtype=torch.int16, device="cuda")
            dec = codec.decode_device(comp, boff, off, M, L, out=dec_buf[shift:shift + n])
            assert np.array_equal(dec.cpu().numpy(), x), (shift, M, L)
            assert int(dec_buf[:shift].abs().sum()) == 0 and int(dec_buf[shift + n:].abs().sum()) == 0


def test_many_chunks_many_waves(codec, oracle):
    """77 chunks x 200 waves of 3500 (Nab-like), M=4: look-back across ~15k waves."""
    from deltarice_b200.synth import nab_like
    x = nab_like(77 * 200, 3500, seed=3).ravel()
    off = np.arange(78, dtype=np.uint64) * (200 * 3500)
    got, boff = codec.encode_host(x, off, 4, 3500)
    for c in (0, 1, 38, 76):
        want = oracle.encode_chunk(x[int(off[c]):int(off[c + 1])], 4, 3500)
        assert np.array_equal(got[int(boff[c]):int(boff[c + 1])].view(np.uint32), want), c
    assert np.array_equal(codec.decode_host(got, boff, off, 4, 3500), x)


def test_errors(codec, oracle):
    import deltarice_b200 as d
    from deltarice_b200 import _lib
    import ctypes as C
    x = np.random.default_rng(1).integers(-32768, 32768, 7000).astype(np.int16)
    off = np.array([0, 7000], dtype=np.uint64)
    # capacity
    out = np.empty(1000, dtype=np.uint8)
    boff = np.zeros(2, dtype=np.uint64)
    u64p = C.POINTER(C.c_uint64)
    rc = codec._L.drice_encode_batch_host(codec._h, x.ctypes.data, off.ctypes.data_as(u64p), 1, 8, 7000,
                                          out.ctypes.data, out.size, boff.ctypes.data_as(u64p))
    assert rc == _lib.E_CAPACITY
    # bad M
    with pytest.raises(d.DeltaRiceError):
        codec.encode_host(x, off, 12, 7000)
    # malformed stream: truncated record, wrong count, wrong expected size
    s = oracle.encode_chunk(x, 8, 7000)
    with pytest.raises(d.DeltaRiceError):
        codec.decode_host(s[:-3].view(np.uint8), None, None, 8, 7000)
    bad = s.copy()
    bad[1] += 5
    with pytest.raises(d.DeltaRiceError):
        codec.decode_host(bad.view(np.uint8), None, None, 8, 7000)
    with pytest.raises(d.DeltaRiceError):
        codec.decode_host(s.view(np.uint8), None, np.array([0, 6999], dtype=np.uint64), 8, 7000)
    # the context stays usable after errors
    assert np.array_equal(codec.decode_host(s.view(np.uint8), None, None, 8, 7000), x)


@pytest.mark.parametrize("M,L,sizes,sigma", [
    (8, 81920, [32 * 81920], 10.0),                 # reference docs/Performance.md:27: nEDM-like chunk
    (8, 500000, [3 * 500000 + 12345], 30.0),        # NOPTREX-like waves, ragged last wave
    (8, None, [1400000], 10.0),                     # the reference's DEFAULT options: the whole chunk is one wave
    (4, 8193, [8193 * 5, 8193 * 2 + 77, 1], 3.0),   # just past the tile kernels' 8192 samples
    (16, 20000, [20000 * 7], 3000.0),               # escape dominated: 25-bit codes, many segments per wave
    (1, 30000, [30000 * 2], 0.7),                   # RiceParameter 1
    (8, None, [100001, 0, 16385, 1, 9000], 10.0),   # whole-chunk waves of odd sizes (every alignment), an empty chunk, one sample
    (4, 8200, [8200 * 640], 3.0),                   # MANY long waves: one CTA per wave (encode_multi_kernel), lane parser
    (2, 16400, [16400 * 3 + 5], 2.0),               # segments that end inside a word, short ragged last wave
])
def test_long_waves(codec, oracle, M, L, sizes, sigma):
    """Waves longer than 8192 samples: in small batches the encode_long_* kernels (several CTAs per wave: sizing,
    scan, packing at the segments' bit offsets) and parse_long_kernel (several CTAs per wave, chained along the
    record); in large ones encode_multi_kernel and the lane parser.  Bit-exact against the oracle, exact round trip."""
    r = np.random.default_rng(77)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    x = np.clip(np.rint(np.cumsum(r.normal(0, sigma, int(off[-1]))) % 3000 - 1500 + r.normal(0, sigma, int(off[-1]))),
                -32768, 32767).astype(np.int16)
    want, wboff = _oracle_batch(oracle, x, off, M, L)
    got, boff = codec.encode_host(x, off, M, L)
    assert np.array_equal(boff, wboff)
    assert np.array_equal(got.view(np.uint32), want)
    assert np.array_equal(codec.decode_host(got, boff, off, M, L), x)
    # the oracle's stream decodes as well (the parser does not rely on our encoder)
    assert np.array_equal(codec.decode_host(want.view(np.uint8), wboff, off, M, L), x)


def _code_bits(x, k):
    """Bits of every sample's code (delta from the previous sample of the wave, src/deltaRice.c:207-228)."""
    d = np.diff(x.astype(np.int64), prepend=0)
    d = ((d + 32768) % 65536) - 32768
    u = np.where(d >= 0, 2 * d, -2 * d - 1)
    q = u >> k
    return np.where(q >= 8, 25, q + k + 1)


@pytest.mark.parametrize("L,total_hint", [(6000, 4), (None, 1)])
def test_last_code_straddles_a_segment(codec, oracle, L, total_hint):
    """parse_long_kernel cuts records into segments (256 words for a handful of waves): a wave whose LAST code starts
    in one segment and ends in the next - which then holds nothing but that code's tail - is a valid record."""
    k, M = 3, 8
    r = np.random.default_rng(11)
    tail = np.clip(np.rint(np.cumsum(r.normal(0, 6, 40000))), -32768, 32767).astype(np.int16)
    cum = np.cumsum(_code_bits(tail, k))
    # last sample i whose code covers a multiple of 8192 bits (256 words) strictly inside it
    cover = np.nonzero((np.concatenate([[0], cum[:-1]]) // 8192) < ((cum - 1) // 8192))[0]
    cover = cover[(cum[cover] % 8192) != 0]
    cases = [int(cover[0]) + 1, int(cover[len(cover) // 2]) + 1] if L is None else [int(cover[0]) + 1]
    for n0 in cases:
        if L is None:
            x = tail[:n0].copy()                                     # the whole chunk is one wave of n0 samples
        else:
            lead = r.normal(0, 6, L * (total_hint - 1)).astype(np.int16)
            x = np.concatenate([lead, tail[:n0]])                    # ... the chunk's short last wave
            assert n0 < L
        off = np.array([0, x.size], dtype=np.uint64)
        want = oracle.encode_chunk(x, M, L)
        assert int(want[-(int(np.ceil(cum[n0 - 1] / 32))) - 1]) == int(np.ceil(cum[n0 - 1] / 32))   # the last record is the crafted wave
        assert (int(np.ceil(cum[n0 - 1] / 32)) - 1) % 256 == 0      # ... and its last word starts a segment
        got, boff = codec.encode_host(x, off, M, L)
        assert np.array_equal(got.view(np.uint32), want)
        assert np.array_equal(codec.decode_host(want.view(np.uint8), None, off, M, L), x)
        # a word too many / too few behind it is still an error
        import deltarice_b200 as d
        extra = np.concatenate([want, np.zeros(1, np.uint32)]); extra[-(int(np.ceil(cum[n0 - 1] / 32))) - 2] += 1
        with pytest.raises(d.DeltaRiceError):
            codec.decode_host(extra.view(np.uint8), None, off, M, L)


def test_mixed_density_batch_sorted_decode(codec, oracle):
    """A batch whose records are heavy on average (escapes) is decoded with the waves sorted by bits per sample
    (wave_hist_kernel / wave_scatter_kernel) and with heavy warps skipping the table: only the ORDER of decoding
    changes.  Noise levels from 1 to 3000 interleaved wave by wave, several chunks, ragged last chunk; a corrupt
    record is still reported."""
    import deltarice_b200 as d
    r = np.random.default_rng(21)
    L, M, nw = 1200, 8, 2600                                          # (> 896 waves: the lane parser)
    sig = np.array([1, 3, 10, 30, 100, 1000, 3000, 2])[np.arange(nw) % 8][:, None]
    x = np.clip(np.rint(r.normal(0, 1, (nw, L)) * sig), -32768, 32767).astype(np.int16).ravel()[:-77]
    off = d.chunk_offsets(300 * L, x.size)
    want, wboff = _oracle_batch(oracle, x, off, M, L)
    assert want.size * 32 > x.size * (3 + 5)                          # dense enough to take the sorted path
    assert np.array_equal(codec.decode_host(want.view(np.uint8), wboff, off, M, L), x)
    got, boff = codec.encode_host(x, off, M, L)
    assert np.array_equal(got.view(np.uint32), want)
    bad = want.copy()
    bad[int(wboff[3]) // 4 + 1] += 1                                  # first record of chunk 3 claims one word too many
    with pytest.raises(d.DeltaRiceError):
        codec.decode_host(bad.view(np.uint8), wboff, off, M, L)
    assert np.array_equal(codec.decode_host(want.view(np.uint8), wboff, off, M, L), x)


def test_long_wave_capacity_and_filter(codec, oracle):
    """The several-CTAs-per-wave encoder: an output buffer that is too small is reported (nothing is written
    past it), and option tuples whose pre-filter is [1] (no delta) take the same kernels."""
    import deltarice_b200 as d
    import torch
    r = np.random.default_rng(9)
    x = np.clip(np.rint(np.cumsum(r.normal(0, 5, 50000)) % 2000 - 1000), -32768, 32767).astype(np.int16)
    off = np.array([0, x.size], dtype=np.uint64)
    want = oracle.encode_chunk(x, 8, 25000)
    xd = torch.from_numpy(x).cuda()
    small = torch.full((want.size * 4 // 2 + 64,), 0xA5, dtype=torch.uint8, device="cuda")
    boff = torch.zeros(2, dtype=torch.int64, device="cuda")
    status = torch.zeros(2, dtype=torch.int32, device="cuda")
    codec.encode_device_async(xd, off, 8, 25000, small[:want.size * 2], boff, status)
    torch.cuda.synchronize()
    assert int(status[0]) & 1                                     # DRICE capacity flag
    assert int(boff[1]) == want.size * 4                          # the needed size is still reported
    assert bool((small[want.size * 2:] == 0xA5).all())            # nothing past the capacity
    got, _ = codec.encode_host(x, off, 8, 25000)
    assert np.array_equal(got.view(np.uint32), want)
    # filter [1]: the samples are coded as they are
    f = d.DeltaRice(0)
    f.set_filter([1])
    xs = (x // 64).astype(np.int16)
    wantf = oracle.encode_chunk(xs, 8, 25000, filt=[1])
    gotf, bo = f.encode_host(xs, off, 8, 25000)
    assert np.array_equal(gotf.view(np.uint32), wantf)
    assert np.array_equal(f.decode_host(gotf, bo, off, 8, 25000), xs)
    f.close()


def test_long_wave_errors(codec, oracle):
    import deltarice_b200 as d
    x = np.random.default_rng(5).normal(0, 20, 3 * 40000).astype(np.int16)
    s = oracle.encode_chunk(x, 8, 40000)
    assert np.array_equal(codec.decode_host(s.view(np.uint8), None, None, 8, 40000), x)
    for bad in (s[:-5], np.concatenate([s[:20000], s[20001:]])):     # truncated; a word missing inside a record
        with pytest.raises(d.DeltaRiceError):
            codec.decode_host(np.ascontiguousarray(bad).view(np.uint8), None, np.array([0, x.size], dtype=np.uint64), 8, 40000)
    flip = s.copy()
    flip[30000] = 0                                                  # a run of 32 zero bits: not a code
    with pytest.raises(d.DeltaRiceError):
        codec.decode_host(flip.view(np.uint8), None, None, 8, 40000)
    assert np.array_equal(codec.decode_host(s.view(np.uint8), None, None, 8, 40000), x)


def test_calls_on_alternating_streams(codec, oracle):
    """One context, calls enqueued on two different streams back to back with no synchronisation in
    between: the context's scratch is shared, so a call must be ordered behind the previous call's
    last kernel (include/deltarice_b200.h).  Results stay bit-exact."""
    import torch
    from deltarice_b200.synth import nab_like
    import deltarice_b200 as d
    L, M, wpc = 3500, 4, 200
    xs = [torch.from_numpy(nab_like(4000, L, seed=s).ravel()).cuda() for s in (1, 2)]
    off = d.chunk_offsets(wpc * L, xs[0].numel())
    outs = [torch.empty(codec.bound_bytes(off, L), dtype=torch.uint8, device="cuda") for _ in xs]
    boffs = [torch.zeros(len(off), dtype=torch.int64, device="cuda") for _ in xs]
    st = [torch.zeros(2, dtype=torch.int32, device="cuda") for _ in range(4)]
    ys = [torch.empty_like(x) for x in xs]
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(5):
        with torch.cuda.stream(sa):
            codec.encode_device_async(xs[0], off, M, L, outs[0], boffs[0], st[0])
        with torch.cuda.stream(sb):
            codec.encode_device_async(xs[1], off, M, L, outs[1], boffs[1], st[1])
    torch.cuda.synchronize()
    hb = [b.cpu().numpy().astype(np.uint64) for b in boffs]
    for rep in range(5):
        with torch.cuda.stream(sa):
            codec.decode_device_async(outs[0][:int(hb[0][-1])], hb[0], off, M, L, ys[0], st[2])
        with torch.cuda.stream(sb):
            codec.decode_device_async(outs[1][:int(hb[1][-1])], hb[1], off, M, L, ys[1], st[3])
    torch.cuda.synchronize()
    assert all(int(s[0]) == 0 for s in st)
    for i in range(2):
        x = xs[i].cpu().numpy()
        assert torch.equal(ys[i], xs[i])
        for c in (0, len(off) - 2):
            want = oracle.encode_chunk(x[int(off[c]):int(off[c + 1])], M, L)
            got = outs[i][int(hb[i][c]):int(hb[i][c + 1])].cpu().numpy().view(np.uint32)
            assert np.array_equal(got, want)


def test_full_size_c2_roundtrip_and_sampled_parity(codec, oracle):
    """BASELINE config C2 at full size (153 391 Nab-like waves of 3500, M=4, ~1 GB): device
    round trip decode(encode(x)) == x, compression ratio, and byte parity with the oracle on
    sampled chunks."""
    import torch
    from deltarice_b200.synth import nab_like_torch
    n_waves, L, M, wpc = 153391, 3500, 4, 2000
    x = nab_like_torch(n_waves, L, 20251018, "cuda").reshape(-1)
    from deltarice_b200 import chunk_offsets
    off = chunk_offsets(wpc * L, x.numel())
    comp, boff = codec.encode_device(x, off, M, L)
    ratio = comp.numel() / (x.numel() * 2)
    assert 0.22 < ratio < 0.32, ratio
    y = codec.decode_device(comp, boff, off, M, L)
    assert torch.equal(x, y)
    for c in (0, 37, len(off) - 2):
        xs = x[int(off[c]):int(off[c + 1])].cpu().numpy()
        want = oracle.encode_chunk(xs, M, L, mt=True)
        got = comp[int(boff[c]):int(boff[c + 1])].cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), c
    # size-independent property: total bytes = 4*(chunks + waves + sum nwords) and every chunk
    # stream starts with its sample count
    heads = comp.view(torch.int32)[torch.from_numpy((boff[:-1] // 4).astype(np.int64)).cuda()]
    assert torch.equal(heads.cpu(), torch.from_numpy(np.diff(off).astype(np.int32)))


@pytest.mark.parametrize("M", [1, 2, 4, 8, 16, 32, 64])
def test_config_c3_rice_parameter_sweep(codec, oracle, M):
    """BASELINE config C3 (RiceParameter sweep 1..64 at WaveformLength 7000): a 128 MiB slice per
    parameter on the device path - round trip, the checksum-of-sizes property (stream bytes =
    4*(chunks + waves + sum nwords)), and byte parity with the oracle on sampled chunks."""
    import torch
    from deltarice_b200 import chunk_offsets
    from deltarice_b200.synth import nab_like_torch
    n_waves, L, wpc = 9587, 7000, 2000
    x = nab_like_torch(n_waves, L, 77 + M, "cuda").reshape(-1)
    off = chunk_offsets(wpc * L, x.numel())
    comp, boff = codec.encode_device(x, off, M, L)
    y = codec.decode_device(comp, boff, off, M, L)
    assert torch.equal(x, y)
    words = comp.view(torch.int32)
    total = 0
    for c in (0, len(off) - 2):
        xs = x[int(off[c]):int(off[c + 1])].cpu().numpy()
        want = oracle.encode_chunk(xs, M, L, mt=True)
        got = comp[int(boff[c]):int(boff[c + 1])].cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), (M, c)
    # walk the headers of the whole stream on the host: sizes must add up exactly
    w = words.cpu().numpy().view(np.uint32)
    for c in range(len(off) - 1):
        cur = int(boff[c]) // 4
        assert w[cur] == int(off[c + 1] - off[c])
        cur += 1
        nw = -(-int(off[c + 1] - off[c]) // L)
        for _ in range(nw):
            cur += int(w[cur]) + 1
        assert cur == int(boff[c + 1]) // 4, (M, c)
        total += cur - int(boff[c]) // 4
    assert total * 4 == comp.numel()


def test_config_c4_mixed_noise_decode(codec, oracle):
    """BASELINE config C4 (decode of a pre-compressed stream with mixed noise levels: long unary
    runs next to escape-dominated waves in the same warp): 64 MiB slice, stream checked against the
    oracle on sampled chunks, then decoded on the device and compared with the input."""
    import torch
    from deltarice_b200 import chunk_offsets
    from deltarice_b200.synth import gaussian_mix
    n_waves, L, M, wpc = 4794, 7000, 8, 600
    xh = gaussian_mix(n_waves, L, seed=11).ravel()
    x = torch.from_numpy(xh).cuda()
    off = chunk_offsets(wpc * L, x.numel())
    comp, boff = codec.encode_device(x, off, M, L)
    for c in (0, 3, len(off) - 2):
        want = oracle.encode_chunk(xh[int(off[c]):int(off[c + 1])], M, L, mt=True)
        got = comp[int(boff[c]):int(boff[c + 1])].cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), c
    y = codec.decode_device(comp, boff, off, M, L)
    assert torch.equal(x, y)
    # decode of the oracle's own stream for one chunk (not produced by our encoder)
    c = 1
    s = oracle.encode_chunk(xh[int(off[c]):int(off[c + 1])], M, L, mt=True)
    back = codec.decode_host(s.view(np.uint8), None, None, M, L)
    assert np.array_equal(back, xh[int(off[c]):int(off[c + 1])])


def test_direct_chunk_batch_front_end(codec, oracle):
    """all chunks of a dataset through ONE batch call; every stored chunk must be
    what the filter produces for that chunk (README config: (100,7000) int16, chunks (20,7000),
    opts (8, 7000)), edge chunks padded as HDF5 pads them, generic opts included."""
    from deltarice_b200 import h5
    data = np.random.default_rng(0).normal(0, 10, (100, 7000)).astype(np.int16)
    stored = h5.encode_dataset_chunks(codec, data, (20, 7000), (8, 7000))
    assert [o for o, _ in stored] == [(i, 0) for i in range(0, 100, 20)]
    for (org, b) in stored:
        want = oracle.encode_chunk(data[org[0]:org[0] + 20].ravel(), 8, 7000)
        assert np.array_equal(np.frombuffer(b, np.uint32), want)
        assert b == h5.apply_filter(data[org[0]:org[0] + 20].tobytes(), (8, 7000))
    back = h5.decode_dataset_chunks(codec, [b for _, b in stored], data.shape, (20, 7000), (8, 7000))
    assert np.array_equal(back, data)
    # ragged grid (edge chunks padded with the fill value) and an option tuple with a pre-filter
    d2 = np.random.default_rng(1).integers(-3000, 3000, (50, 333)).astype(np.int16)
    opts = (4, 0xFFFFFFFF, 3, 1, 0xFFFFFFFE, 1)
    stored = h5.encode_dataset_chunks(codec, d2, (16, 128), opts)
    assert len(stored) == 4 * 3
    org, b = stored[-1]
    block = np.zeros((16, 128), np.int16)
    block[:50 - org[0], :333 - org[1]] = d2[org[0]:, org[1]:]
    assert np.array_equal(np.frombuffer(b, np.uint32), oracle.encode_chunk(block.ravel(), 4, None, filt=[1, -2, 1]))
    back = h5.decode_dataset_chunks(codec, [bb for _, bb in stored], d2.shape, (16, 128), opts)
    assert np.array_equal(back, d2)


def test_more_than_65535_chunks(codec, oracle):
    """Grids are sliced at 65535 chunks in y (header scan, pre-filter): 70 000 one-wave chunks of
    4096 samples on the device path, with and without a pre-filter; sampled chunks against the oracle."""
    import torch
    from deltarice_b200 import chunk_offsets
    n_chunks, n, M = 70000, 4096, 4
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = (torch.randn(n_chunks * n, generator=g, device="cuda") * 6).cumsum(0).remainder(4000).to(torch.int16)
    off = chunk_offsets(n, x.numel())
    for taps in (None, [1, -2, 1]):
        codec.set_filter(taps)
        try:
            comp, boff = codec.encode_device(x, off, M, n)
            y = codec.decode_device(comp, boff, off, M, n)
        finally:
            codec.set_filter(None)
        assert torch.equal(x, y), taps
        for c in (0, 65534, 65535, 65536, n_chunks - 1):
            xs = x[int(off[c]):int(off[c + 1])].cpu().numpy()
            want = oracle.encode_chunk(xs, M, n, filt=taps)
            got = comp[int(boff[c]):int(boff[c + 1])].cpu().numpy().view(np.uint32)
            assert np.array_equal(got, want), (taps, c)
