"""Generates tests/golden/golden_filters_v1.npz: streams AND decoded outputs of the UNMODIFIED
reference (oracle/_ref, built by `make -C oracle ref`; needs /root/reference, i.e. runs in the
build container only) for compression_opts that carry a pre-filter (cd_nelmts >= 3, reference
src/deltaRice.c:277-290, encode :64-74, decode :91-102).  The decoded output is stored too:
filters whose first tap is not +-1 do not round-trip (the reference divides by it).

    python tests/golden/make_golden_filters.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def cases():
    r = np.random.default_rng(32025)
    u16 = r.uniform(-32768, 32768, size=2 ** 14).astype(np.int16)
    smooth = np.cumsum(r.normal(0, 6, 7000 * 2)).astype(np.int16)
    out = []
    # reference tests/test.py:46-83: compression_opts (8, 1024, 1, 1) = "no filter"
    out.append(("ref_test_different_filter", u16, O.cd_values(8, 1024, [1])))
    out.append(("ref_test_all_signed", np.arange(-32768, 32768).astype(np.int16), O.cd_values(8, 1024, [1])))
    out.append(("ref_test_all_unsigned", np.arange(0, 65536).astype(np.uint16).view(np.int16), O.cd_values(8, 1024, [1])))
    out.append(("identity_whole_chunk", r.normal(0, 30, 5000).astype(np.int16), O.cd_values(16, None, [1])))
    out.append(("second_difference", smooth, O.cd_values(4, 7000, [1, -2, 1])))
    out.append(("second_difference_leftover", smooth[:7000 + 321], O.cd_values(4, 3500, [1, -2, 1])))
    out.append(("negated_delta", smooth, O.cd_values(8, 3500, [-1, 1])))
    out.append(("sparse_taps", r.normal(0, 40, 9000).astype(np.int16), O.cd_values(32, 3000, [1, 0, 0, -1])))
    out.append(("four_taps_wrapping", u16, O.cd_values(8, 1024, [3, 1, -2, 5])))
    out.append(("lossy_first_tap_2", r.normal(0, 50, 4096).astype(np.int16), O.cd_values(8, 512, [2, -1])))
    out.append(("lossy_first_tap_m3", r.normal(0, 50, 4096).astype(np.int16), O.cd_values(8, 512, [-3, 7])))
    out.append(("sixteen_taps", r.normal(0, 10, 6000).astype(np.int16), O.cd_values(8, 2000, [1] + [(-1) ** i * (i % 3) for i in range(1, 16)])))
    out.append(("taps_longer_than_wave", r.normal(0, 10, 40).astype(np.int16), O.cd_values(8, 3, [1, -1, 1, -1, 1])))
    out.append(("identity_long_wave", r.normal(0, 25, 30000).astype(np.int16), O.cd_values(8, 20000, [1])))
    out.append(("generic_long_wave", np.cumsum(r.normal(0, 4, 30000)).astype(np.int16), O.cd_values(8, 20000, [1, -2, 1])))
    return out


def main():
    assert O.ref_available("omp"), "build oracle/_ref first: make -C oracle ref"
    blob, names = {}, []
    for name, x, cd in cases():
        stream = np.frombuffer(O.ref_filter(x, cd, reverse=False), dtype=np.uint32)
        back = np.frombuffer(O.ref_filter(stream, cd, reverse=True), dtype=np.int16)
        blob[f"{name}__x"] = x
        blob[f"{name}__cd"] = np.array(cd, dtype=np.uint32)
        blob[f"{name}__stream"] = stream
        blob[f"{name}__back"] = back
        names.append(name)
    blob["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_filters_v1.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes,", len(names), "vectors")


if __name__ == "__main__":
    main()
