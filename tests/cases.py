"""Shared seeded parity cases: (name, samples int16[total], M, L) — sizes the oracle finishes
in milliseconds.  Covers the reference's own test inputs (reference tests/test.py) and the
edge cases of SURVEY.md §7 "bit-exact corner cases"."""
import numpy as np


def _rng(seed):
    return np.random.default_rng(seed)


def small_cases():
    cs = []
    r = _rng(0)
    # README / C1 chunk: (20,7000) N(0,10), M=8, L=7000  (reference README.md:71-91)
    cs.append(("readme_chunk", r.normal(0, 10, (20, 7000)).astype(np.int16).ravel(), 8, 7000))
    # reference tests/test.py: uniform random 2^16 samples; default opts, (16,), (8,1024)
    u = _rng(1).uniform(-32768, 32768, size=2 ** 16).astype(np.int16)
    cs.append(("worst_case_default", u, 8, None))
    cs.append(("different_m", u, 16, None))
    cs.append(("different_m_segment", u, 8, 1024))
    # brute force: every int16 value / every uint16 bit pattern (tests/test.py:59-83)
    cs.append(("all_signed", np.arange(-32768, 32768).astype(np.int16), 8, 1024))
    cs.append(("all_unsigned", np.arange(0, 65536).astype(np.uint16).view(np.int16), 8, 1024))
    # doc example deltas [-2, 25, 0...] (docs/Algorithm.md:9)
    d = np.array([-2, 25] + [0] * 30, dtype=np.int16)
    cs.append(("doc_example", np.cumsum(d).astype(np.int16), 8, 32))
    cs.append(("zeros", np.zeros(32, np.int16), 8, None))
    cs.append(("alternating_extremes", np.array([-32768, 32767] * 16, dtype=np.int16), 8, 32))
    # leftover (short last wave), tiny waves, single sample
    cs.append(("leftover", _rng(2).normal(0, 20, 7000 * 3 + 123).astype(np.int16), 8, 7000))
    cs.append(("leftover_small", _rng(3).normal(0, 5, 1000).astype(np.int16), 4, 33))
    cs.append(("one_sample", np.array([-7], dtype=np.int16), 8, None))
    cs.append(("one_sample_waves", _rng(4).normal(0, 300, 50).astype(np.int16), 2, 1))
    cs.append(("L_gt_total", _rng(5).normal(0, 9, 100).astype(np.int16), 8, 7000))
    # every legal M; M=1 only with small deltas (reference hangs otherwise, Appendix B3)
    base = _rng(6).normal(0, 40, 3500 * 4).astype(np.int16)
    for k in range(1, 16):
        cs.append((f"M_2^{k}", base, 1 << k, 3500))
    cs.append(("M_1_small", _rng(7).normal(0, 1.5, 3500 * 2).astype(np.int16), 1, 3500))
    cs.append(("M_1_const", np.full(4096, 1234, np.int16), 1, 1024))      # 1 bit/sample: many threads per word
    cs.append(("M_2_const", np.full(5000, -5, np.int16), 2, 2500))
    # bits%32 == 0 exactly: 8 samples of 4 bits (M=8, delta 0) -> no pad word
    cs.append(("exact_word", np.zeros(16, np.int16), 8, 8))
    # escape heavy / all escapes
    cs.append(("all_escape", _rng(8).integers(-32768, 32768, 7000 * 2).astype(np.int16), 2, 7000))
    # long waves: single-tile limit and beyond (multi-tile encoder)
    long = _rng(9).normal(0, 12, 8177 * 2 + 5).astype(np.int16)
    cs.append(("L_8177", long, 8, 8177))
    cs.append(("L_8178", long, 8, 8178))
    cs.append(("L_20000", _rng(10).normal(0, 30, 50001).astype(np.int16), 4, 20000))
    cs.append(("L_81920", _rng(11).normal(0, 6, 81920 * 2).astype(np.int16), 8, 81920))
    cs.append(("whole_chunk_200k", _rng(12).normal(0, 100, 200001).astype(np.int16), 16, None))
    # odd wave lengths (unaligned wave starts)
    cs.append(("L_odd_3501", _rng(13).normal(0, 10, 3501 * 5).astype(np.int16), 4, 3501))
    cs.append(("L_odd_7", _rng(14).normal(0, 10, 7 * 100 + 3).astype(np.int16), 8, 7))
    cs.append(("L_3500_nab", None, 4, 3500))   # filled below
    from deltarice_b200.synth import nab_like, gaussian_mix
    cs[-1] = ("L_3500_nab", nab_like(64, 3500, seed=1).ravel(), 4, 3500)
    cs.append(("L_7000_mix", gaussian_mix(24, 7000, seed=2).ravel(), 8, 7000))
    return cs
