/* drice_oracle.c — CPU restatement of the Delta-Rice stream format.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under deltarice_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement byte-for-byte
 * against (1) the three known-answer streams recorded in SURVEY.md §8c, (2) the prose
 * example of reference docs/Algorithm.md:9, (3) golden vectors produced by the
 * UNMODIFIED reference src/deltaRice.c compiled into oracle/_ref/ (tests/golden/,
 * generators tests/golden/make_golden.py and make_golden_filters.py: the latter pins the
 * generic pre-filter branches, streams and decoded outputs), and (4) oracle/_ref itself
 * when present.
 *
 * Every function cites the reference lines (relative to /root/reference/) it follows.
 * The code is written from the format description, not transcribed.
 *
 * Defined-domain decisions where the reference is undefined:
 *   - M must be 2^k, 0 <= k <= 15.  (B4/B5: reference corrupts or emits garbage.)
 *   - k == 0 and zig-zag value >= 32768: reference hangs (B3); here the Appendix-A rule
 *     is applied uniformly (quotient >= 8 -> escape).
 *   - leftover (short last wave) follows the reference's OpenMP build (B7).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define DRICE_ESC_Q 8 /* "giveup" quotient, reference src/deltaRice.c:203 */

/* reference src/deltaRice.c:114-136 (determinePowerOf2), restricted to the defined domain */
int drice_oracle_log2_param(int M)
{
    if (M <= 0 || (M & (M - 1)) != 0) return -1;
    int k = 0;
    while ((1 << k) != M) ++k;
    return k <= 15 ? k : -1;
}

/* worst case: every sample escapes to 25 bits (reference src/deltaRice.c:223-228) */
size_t drice_oracle_wave_bound_words(size_t nsamples)
{
    return (25 * nsamples + 31) / 32;
}

/* number of waves of a chunk, reference src/deltaRice.c:399-403 */
size_t drice_oracle_num_waves(size_t total, size_t L)
{
    return L ? (total + L - 1) / L : 0;
}

size_t drice_oracle_chunk_bound_words(size_t total, size_t L)
{
    if (L == 0 || L > total) L = total;
    if (total == 0) return 1;
    size_t W = drice_oracle_num_waves(total, L);
    size_t tail = total - (W - 1) * L;
    return 1 + W + (W - 1) * drice_oracle_wave_bound_words(L) + drice_oracle_wave_bound_words(tail);
}

/* zig-zag of a wrapped 16-bit delta, reference src/deltaRice.c:207-211 */
static inline uint32_t zigzag16(int16_t d)
{
    int32_t v = (int32_t)d * 2;
    return v >= 0 ? (uint32_t)v : (uint32_t)(-v - 1);
}

/* inverse, reference src/deltaRice.c:172-177 */
static inline int16_t unzigzag16(uint32_t u)
{
    return (u & 1u) ? (int16_t)(-(int32_t)((u + 1) >> 1)) : (int16_t)(u >> 1);
}

/* Pre-filter value of sample i of a wave (reference src/deltaRice.c:49-75, encodeWaveform):
 * the delta filter [1,-1] (f == NULL here) is d[0] = x[0], d[i] = x[i] - x[i-1] (:53-62); any
 * other filter is the FIR sum over the taps that stay inside the wave, accumulated in a
 * `short` (:64-74), i.e. modulo 2^16. */
static inline int16_t prefilter_at(const int16_t *x, size_t i, const int *f, int flen)
{
    if (!f) return (i == 0) ? x[0] : (int16_t)(uint16_t)((uint16_t)x[i] - (uint16_t)x[i - 1]);
    uint32_t acc = (uint32_t)((int32_t)x[i] * f[0]);
    for (int j = 1; j < flen && (size_t)j <= i; ++j) acc += (uint32_t)((int32_t)x[i - (size_t)j] * f[j]);
    return (int16_t)(uint16_t)acc;
}

/* is (f, flen) the delta filter?  reference src/deltaRice.c:38-46 (checkIfDeltaFilter) */
static inline int is_delta_filter(const int *f, int flen) { return !f || (flen == 2 && f[0] == 1 && f[1] == -1); }

/* One wave: pre-filter + Rice pack (:205-241).
 * Writes the code words to out[0..], returns the word count.  MSB-first packing:
 * a 64-bit shift register is drained one 32-bit word at a time (:229-235); the last
 * partial word is left-aligned with zero fill (:237-241). */
static size_t encode_wave_f(const int16_t *x, size_t n, int k, const int *f, int flen, uint32_t *out)
{
    uint64_t acc = 0;  /* low `fill` bits are pending output */
    unsigned fill = 0;
    size_t nw = 0;
    if (is_delta_filter(f, flen)) f = NULL;
    for (size_t i = 0; i < n; ++i) {
        int16_t d = prefilter_at(x, i, f, flen);
        uint32_t u = zigzag16(d);
        uint32_t q = u >> k;
        if (q < DRICE_ESC_Q) {                     /* q zeros, a one, k remainder bits */
            acc = (acc << (q + 1 + (unsigned)k)) | (1u << k) | (u & ((1u << k) - 1u));
            fill += q + 1 + (unsigned)k;
        } else {                                   /* 8 zeros, a one, zig-zag value in 16 bits */
            acc = (acc << 25) | (1u << 16) | u;
            fill += 25;
        }
        if (fill >= 32) {
            out[nw++] = (uint32_t)(acc >> (fill - 32));
            fill -= 32;
            acc &= (fill ? ((1ull << fill) - 1ull) : 0ull);
        }
    }
    if (fill) out[nw++] = (uint32_t)(acc << (32 - fill));
    return nw;
}

size_t drice_oracle_encode_wave(const int16_t *x, size_t n, int k, uint32_t *out)
{
    return encode_wave_f(x, n, k, NULL, 0, out);
}

/* Chunk framing, reference src/deltaRice.c:383-436 (OpenMP branch semantics):
 *   out[0] = total samples; then per wave [nwords][words...] (:379,:415,:427-432).
 * L == 0 means "whole chunk is one wave" (WaveformLength -1, :391-393).
 * Returns words written, or 0 on error (bad M, capacity). */
size_t drice_oracle_encode_chunk_f(const int16_t *x, size_t total, int M, size_t L, const int *f, int flen,
                                   uint32_t *out, size_t cap_words)
{
    int k = drice_oracle_log2_param(M);
    if (k < 0 || total > 0x7fffffffu) return 0;
    if (f && (flen < 1 || f[0] == 0)) return 0;      /* empty filter / division by f[0] (:100), Appendix B9 */
    if (L == 0) L = total;
    if (cap_words < drice_oracle_chunk_bound_words(total, L)) return 0;
    out[0] = (uint32_t)total;
    size_t pos = 1;
    for (size_t s = 0; s < total; s += L) {
        size_t n = total - s < L ? total - s : L;     /* short last wave, :420-422 */
        size_t nw = encode_wave_f(x + s, n, k, f, flen, out + pos + 1);
        out[pos] = (uint32_t)nw;
        pos += nw + 1;
    }
    return pos;
}

size_t drice_oracle_encode_chunk(const int16_t *x, size_t total, int M, size_t L,
                                 uint32_t *out, size_t cap_words)
{
    return drice_oracle_encode_chunk_f(x, total, M, L, NULL, 0, out, cap_words);
}

/* One wave: Rice parse (reference src/deltaRice.c:154-187) + inverse delta (:80-89).
 * `avail` = words available from `in` (bounds guard, reference has none: Appendix B8).
 * Returns the number of words consumed by the codes, or (size_t)-1 on a malformed
 * stream (unary run longer than 8 or reading past `avail`). */
static size_t decode_wave_f(const uint32_t *in, size_t avail, size_t n, int k, const int *f, int flen, int16_t *y)
{
    uint64_t pos = 0; /* bit position */
    int16_t acc = 0;
    if (is_delta_filter(f, flen)) f = NULL;
    for (size_t i = 0; i < n; ++i) {
        unsigned q = 0;
        for (;;) {                                /* unary run, :156-159 */
            if ((pos >> 5) >= avail) return (size_t)-1;
            unsigned bit = (in[pos >> 5] >> (31 - (pos & 31))) & 1u;
            ++pos;
            if (bit) break;
            if (++q > DRICE_ESC_Q) return (size_t)-1;
        }
        unsigned nb = (q == DRICE_ESC_Q) ? 16u : (unsigned)k;  /* :161-171 */
        uint32_t v = 0;
        for (unsigned b = 0; b < nb; ++b) {
            if ((pos >> 5) >= avail) return (size_t)-1;
            v = (v << 1) | ((in[pos >> 5] >> (31 - (pos & 31))) & 1u);
            ++pos;
        }
        uint32_t u = (q == DRICE_ESC_Q) ? v : ((q << k) + v);
        int16_t d = unzigzag16(u);
        if (!f) {
            acc = (i == 0) ? d : (int16_t)(uint16_t)((uint16_t)acc + (uint16_t)d);
        } else {
            /* inverse of a generic filter, reference src/deltaRice.c:91-102 (decodeWaveform): the
             * recursion runs in a `short` (modulo 2^16), then a C division by f[0] */
            uint32_t t = (uint16_t)d;
            for (int j = 1; j < flen && (size_t)j <= i; ++j) t -= (uint32_t)((int32_t)y[i - (size_t)j] * f[j]);
            acc = (int16_t)((int32_t)(int16_t)(uint16_t)t / f[0]);
        }
        y[i] = acc;
    }
    return (size_t)((pos + 31) >> 5);
}

size_t drice_oracle_decode_wave(const uint32_t *in, size_t avail, size_t n, int k, int16_t *y)
{
    return decode_wave_f(in, avail, n, k, NULL, 0, y);
}

/* Chunk decode, reference src/deltaRice.c:301-341 (OpenMP branch): total = in[0] (:306),
 * header walk cur += in[cur]+1 (:319-325), waves of L samples, last one short (:329-331).
 * Returns samples written, or (size_t)-1 on error. */
size_t drice_oracle_decode_chunk_f(const uint32_t *in, size_t nwords, int M, size_t L, const int *f, int flen,
                                   int16_t *y, size_t cap_samples)
{
    int k = drice_oracle_log2_param(M);
    if (k < 0 || nwords < 1) return (size_t)-1;
    if (f && (flen < 1 || f[0] == 0)) return (size_t)-1;
    size_t total = in[0];
    if (total > cap_samples) return (size_t)-1;
    if (L == 0) L = total;
    size_t cur = 1;
    for (size_t s = 0; s < total; s += L) {
        size_t n = total - s < L ? total - s : L;
        if (cur >= nwords) return (size_t)-1;
        size_t nw = in[cur];
        if (cur + 1 + nw > nwords) return (size_t)-1;
        size_t used = decode_wave_f(in + cur + 1, nw, n, k, f, flen, y + s);
        if (used == (size_t)-1 || used != nw) return (size_t)-1;
        cur += nw + 1;
    }
    return total;
}

size_t drice_oracle_decode_chunk(const uint32_t *in, size_t nwords, int M, size_t L,
                                 int16_t *y, size_t cap_samples)
{
    return drice_oracle_decode_chunk_f(in, nwords, M, L, NULL, 0, y, cap_samples);
}

/* ---- multi-chunk helpers for the timed CPU baseline ("port" kind) ----------------
 * Encode/decode `nchunks` equal chunks back to back with OpenMP over WAVES inside each
 * chunk, mirroring the reference's parallel structure (:417, :327). */
#if defined(_OPENMP)
#include <omp.h>
#endif
#include <stdlib.h>

size_t drice_oracle_encode_chunk_mt(const int16_t *x, size_t total, int M, size_t L,
                                    uint32_t *out, size_t cap_words)
{
    int k = drice_oracle_log2_param(M);
    if (k < 0 || total > 0x7fffffffu) return 0;
    if (L == 0) L = total;
    if (total == 0) { if (cap_words < 1) return 0; out[0] = 0; return 1; }
    if (cap_words < drice_oracle_chunk_bound_words(total, L)) return 0;
    size_t W = drice_oracle_num_waves(total, L);
    size_t slot = drice_oracle_wave_bound_words(L) + 1;
    uint32_t *stage = (uint32_t *)malloc(W * slot * sizeof(uint32_t));
    size_t *sizes = (size_t *)malloc(W * sizeof(size_t));
    if (!stage || !sizes) { free(stage); free(sizes); return 0; }
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)W; ++i) {
        size_t s = (size_t)i * L;
        size_t n = total - s < L ? total - s : L;
        size_t nw = drice_oracle_encode_wave(x + s, n, k, stage + (size_t)i * slot + 1);
        stage[(size_t)i * slot] = (uint32_t)nw;
        sizes[i] = nw + 1;
    }
    out[0] = (uint32_t)total;
    size_t pos = 1;
    for (size_t w = 0; w < W; ++w) {
        memcpy(out + pos, stage + w * slot, sizes[w] * sizeof(uint32_t));
        pos += sizes[w];
    }
    free(stage); free(sizes);
    return pos;
}

size_t drice_oracle_decode_chunk_mt(const uint32_t *in, size_t nwords, int M, size_t L,
                                    int16_t *y, size_t cap_samples)
{
    int k = drice_oracle_log2_param(M);
    if (k < 0 || nwords < 1) return (size_t)-1;
    size_t total = in[0];
    if (total > cap_samples) return (size_t)-1;
    if (total == 0) return 0;
    if (L == 0) L = total;
    size_t W = drice_oracle_num_waves(total, L);
    size_t *starts = (size_t *)malloc(W * sizeof(size_t));
    if (!starts) return (size_t)-1;
    size_t cur = 1;
    for (size_t w = 0; w < W; ++w) {
        if (cur >= nwords || cur + 1 + in[cur] > nwords) { free(starts); return (size_t)-1; }
        starts[w] = cur;
        cur += in[cur] + 1;
    }
    int bad = 0;
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)W; ++i) {
        size_t s = (size_t)i * L;
        size_t n = total - s < L ? total - s : L;
        size_t nw = in[starts[i]];
        size_t used = drice_oracle_decode_wave(in + starts[i] + 1, nw, n, k, y + s);
        if (used != nw) {
#pragma omp atomic write
            bad = 1;
        }
    }
    free(starts);
    return bad ? (size_t)-1 : total;
}
