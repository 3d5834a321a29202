"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference compiled into
oracle/_ref (make -C oracle ref; needs /root/reference, i.e. runs in the build container
only).  The .npz travels with the repo; tests compare the oracle restatement (CPU suite) and
the CUDA path (-m gpu) against it.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402


def cases():
    r = np.random.default_rng(2025)
    out = []
    d = np.array([-2, 25] + [0] * 30, dtype=np.int16)
    out.append(("doc_example", np.cumsum(d).astype(np.int16), (8, 32)))
    out.append(("zeros_default", np.zeros(32, np.int16), ()))
    out.append(("alternating_extremes", np.array([-32768, 32767] * 16, dtype=np.int16), (8, 32)))
    out.append(("readme_20x7000", np.random.default_rng(0).normal(0, 10, (20, 7000)).astype(np.int16).ravel(), (8, 7000)))
    out.append(("leftover", r.normal(0, 20, 3500 * 3 + 77).astype(np.int16), (4, 3500)))
    out.append(("uniform_m16", r.uniform(-32768, 32768, 4096).astype(np.int16), (16,)))
    out.append(("uniform_8_1024", r.uniform(-32768, 32768, 8192).astype(np.int16), (8, 1024)))
    out.append(("explicit_delta_filter", r.normal(0, 10, 4096).astype(np.int16), (8, 1024, 2, 1, 0xFFFFFFFF)))
    out.append(("all_signed", np.arange(-32768, 32768).astype(np.int16), (8, 1024)))
    for k in range(1, 11):
        out.append((f"M{1 << k}", r.normal(0, 3 * 2 ** (k / 2), 7000 * 2).astype(np.int16), (1 << k, 7000)))
    out.append(("M1_small", r.normal(0, 1.2, 2048).astype(np.int16), (1, 512)))
    out.append(("escape_heavy", r.integers(-32768, 32768, 7000).astype(np.int16), (2, 3500)))
    out.append(("exact_word", np.zeros(16, np.int16), (8, 8)))
    out.append(("long_wave_20000", r.normal(0, 25, 40000).astype(np.int16), (8, 20000)))
    # round 2: the shapes the several-CTAs-per-wave kernels and the density sort take
    out.append(("default_opts_whole_chunk", r.normal(0, 30, 50001).astype(np.int16), ()))          # RiceParameter 8, one wave
    out.append(("whole_chunk_m4", np.cumsum(r.normal(0, 3, 120001)).astype(np.int16), (4,)))
    out.append(("long_wave_81920", r.normal(0, 6, 81920 * 2).astype(np.int16), (8, 81920)))
    out.append(("long_wave_ragged", r.normal(0, 200, 30000 * 2 + 12345).astype(np.int16), (16, 30000)))
    sig = np.array([1, 3, 10, 30, 100, 1000, 3000, 2])[np.arange(40) % 8][:, None]
    out.append(("mixed_density", np.clip(np.rint(r.normal(0, 1, (40, 1200)) * sig), -32768, 32767).astype(np.int16).ravel(), (8, 1200)))
    # a short last wave whose LAST code starts in the record's 256th word and ends in its 257th
    tail = np.clip(np.rint(np.cumsum(np.random.default_rng(11).normal(0, 6, 40000))), -32768, 32767).astype(np.int16)[:1728]
    out.append(("last_code_straddles_256_words", np.concatenate([r.normal(0, 6, 6000 * 3).astype(np.int16), tail]), (8, 6000)))
    return out


def main():
    assert O.ref_available("omp"), "build oracle/_ref first: make -C oracle ref"
    blob = {}
    names = []
    for name, x, cd in cases():
        stream = np.frombuffer(O.ref_filter(x, cd, reverse=False), dtype=np.uint32)
        back = np.frombuffer(O.ref_filter(stream, cd, reverse=True), dtype=np.int16)
        assert np.array_equal(back, x), name
        blob[f"{name}__x"] = x
        blob[f"{name}__cd"] = np.array(cd, dtype=np.uint32)
        blob[f"{name}__stream"] = stream
        names.append(name)
    blob["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes,", len(names), "vectors")


if __name__ == "__main__":
    main()
