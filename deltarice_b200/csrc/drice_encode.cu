// drice_encode.cu — Delta + Rice ENCODE for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   encodeWaveform delta branch        src/deltaRice.c:49-63
//   compressWithRiceCoding             src/deltaRice.c:191-244
//   perWaveCompression                 src/deltaRice.c:365-381
//   writeWholeCompressedByteString     src/deltaRice.c:383-436 (framing + compaction)
//
// The codec is integer-ALU bound on B200 long before it is HBM bound (the first version of
// this file issued ~61 instructions per sample and ran at 15 % of the HBM roofline), so the
// design goal of the tile kernel below is instructions per sample, not bytes:
//
//   * one CTA per wave ("waveform" of L <= 8192 samples), persistent CTAs taking waves from
//     an atomic ticket, single pass over HBM;
//   * a thread owns 16 consecutive samples = 8 packed int16x2 words.  Delta, zig-zag and the
//     Rice split run on the packed words (PRMT / VIADD.16x2 / LOP3), two samples per
//     instruction where the ISA allows;
//   * two samples are merged into one "pair" code (<= 30 bits when k <= 7) with a single
//     IMAD: value_lo * 2^len_hi + value_hi.  Pairs that contain an escape (quotient >= 8,
//     25-bit code) are rare and take a divergent slow path;
//   * bit offsets: warp shuffle scan + one REDUX over the warp totals (one barrier);
//   * packing appends pairs to a 64/96-bit window with IMAD.WIDE (acc*2^len + value: the
//     multiply IS the shift, and it runs on the FMA pipe, off the saturated ALU pipe) and
//     emits finished 32-bit words to shared memory.  Every thread starts its window with the
//     trailing bits of its predecessor (one shuffle; computed with a 32-bit IMAD chain
//     before offsets are known), so every word is written exactly once, complete — no
//     shared-memory atomics, no fix-up pass;
//   * cross-wave offsets by decoupled look-back over one 64-bit status word per wave; the
//     record [nwords][words] is then copied out coalesced, the first wave of a chunk also
//     writes the chunk header [total].
//
// Waves longer than one tile (L > 8192) use encode_multi_kernel: a sizing sweep, then a
// packing sweep that re-reads the wave and streams completed words straight to HBM.
#include "drice_kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace drice {

namespace {

constexpr uint64_t kFlagAggregate = 1ull << 62;
constexpr uint64_t kFlagPrefix    = 2ull << 62;
constexpr uint64_t kValueMask     = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// 1 << s through PTX so the compiler keeps the multiply form of the appends
__device__ __forceinline__ uint32_t pow2(uint32_t s)
{
    uint32_t d;
    asm("shl.b32 %0, 1, %1;" : "=r"(d) : "r"(s));
    return d;
}

struct WaveGeom {
    uint64_t begin;     // first sample of the wave in raw
    uint32_t n;         // samples in the wave
    uint32_t chunk;     // chunk index
    uint32_t first;     // 1 if first wave of its chunk
    uint32_t chunk_total;
    uint32_t g;         // wave index in the batch
    uint32_t pad_;
};

__device__ __forceinline__ WaveGeom locate_wave(const EncodeParams &p, uint32_t g)
{
    uint32_t c, i;
    if (p.uniform_wpc) {
        c = g / p.uniform_wpc;
        i = g - c * p.uniform_wpc;
    } else {
        uint32_t lo = 0, hi = p.nchunks;   // largest c with chunk_wave_off[c] <= g
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(p.chunk_wave_off + mid) <= g) lo = mid; else hi = mid;
        }
        c = lo;
        i = g - __ldg(p.chunk_wave_off + c);
    }
    const uint64_t cb = __ldg(p.chunk_sample_off + c), ce = __ldg(p.chunk_sample_off + c + 1);
    const uint64_t Lw = p.L ? (uint64_t)p.L : (ce - cb);
    WaveGeom w;
    w.begin = cb + (uint64_t)i * Lw;
    const uint64_t rem = ce - w.begin;
    w.n = (uint32_t)(rem < Lw ? rem : Lw);
    w.chunk = c;
    w.first = (i == 0);
    w.chunk_total = (uint32_t)(ce - cb);
    w.g = g;
    w.pad_ = 0;
    return w;
}

// ======================================================================================
// warp kernel: one WARP per wave (L <= kEncTileMaxL), persistent warps, no block barriers
// ======================================================================================
constexpr int S = kSamplesPerThread;       // 16 samples = 8 packed words per lane and round
constexpr int kRound = 32 * S;             // samples per warp round

// (value, length) of one zig-zag value, escape aware (src/deltaRice.c:212-228)
template <int K>
__device__ __forceinline__ void rice_code(uint32_t u, uint32_t &val, uint32_t &len)
{
    constexpr uint32_t M = 1u << K;
    const uint32_t q = u >> K;
    len = q + (K + 1);
    val = (u & (M - 1u)) | M;
    if (q >= kEscapeQuotient) {
        len = kEscapeBits;
        val = u | 0x10000u;
    }
}

// window state of the packer: `lo` holds the pending bits in its low `n` (< 32) bits; any
// bits above them are stale and never looked at (words are cut out with funnel shifts)
// a register holding a compile-time constant the compiler cannot see through: keeps
// (u | c1) & c2 one three-register LOP3 instead of two immediate forms
__device__ __forceinline__ uint32_t opaque(uint32_t c)
{
    uint32_t r;
    asm("mov.u32 %0, %1;" : "=r"(r) : "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul_wide(uint32_t a, uint32_t b)
{
    uint64_t d;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
    return d;
}

// Bit packer of one lane.  `lo` holds the pending bits in its low `n` (< 32) bits; bits above
// them are stale and never looked at (finished words are cut out with funnel shifts).
// Appending a code is lo * 2^len | value in a 64-bit window: the multiply is the shift and
// runs on the FMA pipe (IMAD.WIDE), off the busier ALU pipe.
template <bool kGuard>
struct Packer {
    uint32_t lo, n;
    uint32_t *ptr;
    uint32_t *end;      // kGuard: nothing is stored at or past `end` (packing straight into HBM)

    __device__ __forceinline__ void store(uint32_t *q, uint32_t v) const
    {
        if (!kGuard || q < end) *q = v;
    }
    // append one code of len <= 31 bits (value < 2^len)
    __device__ __forceinline__ void put(uint32_t v, uint32_t len)
    {
        const uint64_t a = mul_wide(lo, pow2(len));
        const uint32_t alo = (uint32_t)a | v;            // the low `len` bits of the product are 0
        n += len;
        if (n >= 32u) {
            n -= 32u;
            store(ptr++, __funnelshift_r(alo, (uint32_t)(a >> 32), n));
        }
        lo = alo;
    }
};

// 16 consecutive samples of one lane as 8 packed words.  `q` = address of the lane's first
// sample; the widest naturally aligned vector load the wave's start allows is used (warp
// uniform `mis` = (address of the wave's first sample mod 16) / 2).
struct RawWords { uint32_t w[8]; };

__device__ __forceinline__ void load_slot(RawWords &r, const int16_t *q, uint32_t mis)
{
    if (mis == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(q)), b = __ldg(reinterpret_cast<const uint4 *>(q) + 1);
        r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
        r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
    } else if (mis == 4) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const uint2 a = __ldg(reinterpret_cast<const uint2 *>(q) + v);
            r.w[2 * v] = a.x; r.w[2 * v + 1] = a.y;
        }
    } else if ((mis & 1) == 0) {
#pragma unroll
        for (int v = 0; v < 8; ++v) r.w[v] = __ldg(reinterpret_cast<const uint32_t *>(q) + v);
    } else {
        // odd sample offset: 4-byte loads one sample below, halves re-paired
        const uint32_t *qa = reinterpret_cast<const uint32_t *>(q - 1);
        uint32_t t[9];
        t[0] = (uint32_t)(uint16_t)q[0] << 16;
#pragma unroll
        for (int v = 1; v < 8; ++v) t[v] = __ldg(qa + v);
        t[8] = (uint32_t)(uint16_t)q[15];
#pragma unroll
        for (int v = 0; v < 8; ++v) r.w[v] = prmt(t[v], t[v + 1], 0x5432);
    }
}
// slot that may be short or leave the batch buffer [.., hi): `nvalid` samples, rest zero
__device__ __forceinline__ void load_slot_tail(RawWords &r, const int16_t *q, uint32_t mis, uint32_t nvalid,
                                               const int16_t *hi)
{
#pragma unroll
    for (int m = 0; m < 8; ++m) r.w[m] = 0;
    if (nvalid == 0) return;
    if (q + S <= hi) { load_slot(r, q, mis); return; }
#pragma unroll
    for (int i = 0; i < S; ++i)
        if (q + i < hi) r.w[i >> 1] |= (uint32_t)(uint16_t)q[i] << (16 * (i & 1));
}

// exclusive word offset of tile g among all tiles: decoupled look-back by one warp, four rows of
// 32 status words in flight per round trip.  The tile's own aggregate is already published.
template <int kSleepNs = 200>
__device__ __forceinline__ uint64_t lookback_excl(uint64_t *lookback, uint32_t g, uint64_t mine, int lane)
{
    uint64_t excl = 0;
    if (g > 0) {
        int64_t idx = (int64_t)g - 1;
        bool done = false;
        while (!done) {
            uint64_t s[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int64_t my = idx - 32 * r - lane;
                s[r] = my >= 0 ? ld_relaxed_u64(lookback + my) : kFlagPrefix;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (done) break;
                const int64_t my = idx - 32 * r - lane;
                uint64_t v = s[r];
                while (true) {
                    const uint32_t pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                    const uint32_t zm = __ballot_sync(0xffffffffu, (v >> 62) == 0);
                    // entries behind the nearest prefix are not needed
                    const uint32_t need = pm ? ((2u << (__ffs(pm) - 1)) - 1u) : 0xffffffffu;
                    if ((zm & need) == 0) {
                        uint64_t x = ((need >> lane) & 1u) ? (v & kValueMask) : 0ull;
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
                        excl += x;
                        done = pm != 0;
                        break;
                    }
                    __nanosleep(kSleepNs);
                    if ((v >> 62) == 0) v = ld_relaxed_u64(lookback + my);
                }
            }
            idx -= 128;
        }
    }
    if (lane == 0) st_relaxed_u64(lookback + g, kFlagPrefix | (excl + mine));
    return excl;
}

template <int K>
struct RiceConst {
    static constexpr bool kPairs = (K >= 1 && K <= 7);   // two samples merge into one code of <= 30 bits
    static constexpr uint32_t M = 1u << K;
    static constexpr uint32_t MM = M * 0x10001u, NN = (2u * M - 1u) * 0x10001u, QM = 0xFFFFu >> K;
    static constexpr uint32_t HM = ((0xFFFFu << ((K + 3) > 16 ? 16 : (K + 3))) & 0xFFFFu) * 0x10001u;   // quotient >= 8
};

// state of one wave's encoding sweep (warp uniform unless noted)
struct SweepState {
    uint32_t base;          // bits packed so far
    uint32_t carry_round;   // pending bits (left aligned) of the previous round's last lane
    uint32_t prev_last;     // last packed word of the previous round's lane 31
    bool     ovf;           // staging overflowed: sizing only from here on
};

// One round = 512 samples = 16 per lane.  kFull: every lane holds 16 valid samples.
template <int K, bool kDirect, bool kFull, bool kDelta>
__device__ __forceinline__ void encode_round(RawWords &cur, uint32_t nvalid, bool last_lane, int lane,
                                             uint32_t *dst, uint32_t cap, SweepState &st)
{   // kDelta: the delta pre-filter (src/deltaRice.c:53-62); false: the samples are coded as they are
    using C = RiceConst<K>;
    constexpr bool kPairs = C::kPairs;
    constexpr uint32_t M = C::M;
    uint32_t (&w)[8] = cur.w;
    // short last slot: repeat the last valid sample; its codes (delta 0) trail the lane's bits
    // and are cut off below
    // (no pre-filter: the padding samples are zero instead, which codes the same K+1 bits)
    if (!kFull && nvalid > 0 && nvalid < (uint32_t)S) {
        const uint32_t sel_hi = kDelta ? 0x1010u : 0x4410u, sel_all = kDelta ? 0x3232u : 0x4444u;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            if (2u * v + 1 == nvalid) w[v] = prmt(w[v], 0, sel_hi);
            if (v > 0 && 2u * v >= nvalid) w[v] = prmt(w[v - 1], 0, sel_all);
        }
    }
    // word holding the sample before this lane's first one in its HIGH half
    uint32_t pw = __shfl_up_sync(0xffffffffu, w[7], 1);
    if (lane == 0) pw = st.prev_last;
    st.prev_last = __shfl_sync(0xffffffffu, w[7], 31);

    // ---- delta + zig-zag on packed halves: D = per-half (x[j] - x[j-1]);  U = (D + D) ^ sign(D)
    // (src/deltaRice.c:57-62, :207-211)
    uint32_t U[8];
    uint32_t uor = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const uint32_t prev = m ? w[m - 1] : pw;
        const uint32_t X = kDelta ? w[m] * 0xFFFF0001u : w[m];      // high half: hi(w) - lo(w)
        const uint32_t Y = kDelta ? w[m] - (prev >> 16) : w[m];     // low half:  lo(w) - hi(prev)
        const uint32_t D = prmt(Y, X, 0x7610);
        const uint32_t Sg = prmt(D, 0, 0xbb99);             // per-half sign mask
        U[m] = __vadd2(D, D) ^ Sg;
        uor |= U[m];
    }

    // ---- Rice codes: items (value, length) kept in registers across the scan ---------------
    // kPairs: item m = samples 2m, 2m+1 merged into one code of <= 30 bits; an item that holds
    // an escape is flagged (bit 7 of its length) and keeps the packed zig-zag values instead.
    // !kPairs: 16 single-sample items.
    constexpr int NI = kPairs ? 8 : S;
    uint32_t iv[NI], il[NI];
    uint32_t T = 0;                     // bits of this lane
    bool any_flag = false;              // warp uniform: some lane holds an escape
    if (kPairs) {
        const uint32_t MMr = opaque(C::MM), NNr = opaque(C::NN);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const uint32_t u2 = U[m];
            const uint32_t V2 = (u2 | MMr) & NNr;
            const uint32_t qhi = u2 >> (16 + K), qlo = (u2 >> K) & C::QM;
            il[m] = qlo + qhi + 2u * (K + 1);
            iv[m] = mad_lo(V2 & 0xFFFFu, (2u * M) << qhi, V2 >> 16);
        }
        T = ((il[0] + il[1]) + (il[2] + il[3])) + ((il[4] + il[5]) + (il[6] + il[7]));
        const bool flagged = (uor & C::HM) != 0u;
        any_flag = __any_sync(0xffffffffu, flagged);
        if (any_flag) {                 // rare: redo the items that hold an escape
            if (flagged) {
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    if (U[m] & C::HM) {
                        uint32_t v0, l0, v1, l1;
                        rice_code<K>(U[m] & 0xFFFFu, v0, l0);
                        rice_code<K>(U[m] >> 16, v1, l1);
                        T += l0 + l1 - il[m];
                        il[m] = (l0 + l1) | 0x80u;
                        iv[m] = U[m];
                    }
                }
            }
        }
        // padding samples of a short last slot were coded as delta 0: K+1 bits each
        if (!kFull && nvalid < (uint32_t)S) T = nvalid ? T - ((uint32_t)S - nvalid) * (K + 1) : 0u;
    } else {
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const uint32_t u = (j & 1) ? (U[j >> 1] >> 16) : (U[j >> 1] & 0xFFFFu);
            rice_code<K>(u, iv[j], il[j]);
            if (!kFull && (uint32_t)j >= nvalid) { iv[j] = 0; il[j] = 0; }
            T += il[j];
        }
    }

    // ---- warp exclusive scan of T ----------------------------------------------------------
    uint32_t inc = T;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    const uint32_t b0 = st.base + inc - T;           // bit offset of this lane in the wave
    st.base += total;
    if (!kDirect && ((st.base + 31u) >> 5) + 16u > cap) st.ovf = true;
    if (st.ovf) return;                              // warp uniform: sizing only

    // ---- pack ----------------------------------------------------------------------------------
    Packer<kDirect> pk;
    pk.n = b0 & 31u;
    uint32_t *const first = dst + (b0 >> 5);
    pk.ptr = first;
    pk.end = dst + cap;
    pk.lo = 0;
    const bool packs = kFull || nvalid > 0;
    if (packs) {
        if (kPairs) {
            if (!any_flag) {
#pragma unroll
                for (int m = 0; m < 8; ++m) pk.put(iv[m], il[m]);
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    if (il[m] & 0x80u) {
                        uint32_t v0, l0, v1, l1;
                        rice_code<K>(iv[m] & 0xFFFFu, v0, l0);
                        rice_code<K>(iv[m] >> 16, v1, l1);
                        pk.put(v0, l0);
                        pk.put(v1, l1);
                    } else {
                        pk.put(iv[m], il[m]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (kFull || il[j]) pk.put(iv[j], il[j]);
        }
    }
    // ---- stitch the lanes: the bits a lane left pending belong to the first word the next
    // lane wrote (or still holds) --------------------------------------------------------------
    uint32_t frag;                                       // pending bits, left aligned
    asm("shl.b32 %0, %1, %2;" : "=r"(frag) : "r"(pk.lo), "r"(32u - pk.n));   // n == 0 -> 0
    if (!packs) frag = 0;
    const bool flushed = pk.ptr != first;                // wrote its first word itself
    constexpr int kStitch = (K == 0) ? 2 : 1;            // 1-bit codes: a lane may hold < 32 bits
#pragma unroll
    for (int e = 0; e < kStitch; ++e) {
        uint32_t from_prev = __shfl_up_sync(0xffffffffu, frag, 1);
        if (lane == 0) from_prev = st.carry_round;
        if (packs) {
            if (flushed) { if (from_prev) *first |= from_prev; }
            else frag |= from_prev;
        }
    }
    st.carry_round = __shfl_sync(0xffffffffu, frag, 31);
    // the wave's last lane owns the final partial word; bits past the wave's end (padding codes
    // of a short slot) are cleared so the word is zero padded (:237-241)
    if (!kFull && last_lane && packs) {
        if (pk.n) pk.store(pk.ptr, frag);
        if (st.base & 31u) dst[st.base >> 5] &= 0xFFFFFFFFu << (32u - (st.base & 31u));
    }
}

// Encoding sweep over one wave.  kDirect = false: packs into the warp's staging of `cap` words;
// when the wave outgrows it, packing stops (sizing continues) and *overflow is set.
// kDirect = true: packs straight into the record in HBM, `cap` = the wave's word count (nothing
// is stored past it).  Returns the wave's bit count.
template <int K, bool kDirect, bool kDelta>
__device__ __forceinline__ uint32_t encode_wave(const int16_t *wave, uint32_t n, const int16_t *raw_hi, int lane,
                                                uint32_t *dst, uint32_t cap, bool *overflow)
{
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(wave) & 15u) >> 1);
    // rounds in which every lane holds 16 samples inside the buffer; then one generic round
    uint32_t nfull = n / kRound;
    const bool has_tail = (n % kRound) != 0;
    if (!has_tail && nfull) {
        // the wave's last lane must close the final word: run the last full round as the tail
        --nfull;
    }
    SweepState st;
    st.base = 0;
    st.carry_round = 0;
    st.prev_last = 0;
    st.ovf = false;
    // the next round's samples are always in flight while the current round is encoded
    const uint32_t tail_s0 = nfull * kRound + lane * S;
    const int32_t tail_rem = (int32_t)n - (int32_t)tail_s0;
    const uint32_t tail_valid = tail_rem >= S ? (uint32_t)S : (tail_rem > 0 ? (uint32_t)tail_rem : 0u);
    const int16_t *q = wave + lane * S;
    RawWords cur;
    if (nfull) load_slot(cur, q, mis); else load_slot_tail(cur, wave + tail_s0, mis, tail_valid, raw_hi);
    for (uint32_t r = 0; r < nfull; ++r) {
        RawWords now = cur;
        q += kRound;
        if (r + 1 < nfull) load_slot(cur, q, mis); else load_slot_tail(cur, wave + tail_s0, mis, tail_valid, raw_hi);
        encode_round<K, kDirect, true, kDelta>(now, S, false, lane, dst, cap, st);
    }
    encode_round<K, kDirect, false, kDelta>(cur, tail_valid, tail_valid > 0 && tail_s0 + S >= n, lane, dst, cap, st);
    __syncwarp();
    *overflow = st.ovf;
    return st.base;
}

// ---- tile kernel ----------------------------------------------------------------------------
// A tile = NW consecutive waves, taken by one persistent CTA of NW worker warps +
// one control warp.  Per iteration a worker encodes ONE wave of the current tile into one of its
// two staging buffers (single sweep over HBM), then copies out the wave it encoded in the
// previous iteration, whose position has been resolved in the meantime:
//   * the worker that finishes its wave last sums the tile's wave sizes and publishes the tile's
//     aggregate at once (nobody ever waits for a size that is already known);
//   * the control warp resolves tile after tile by decoupled look-back and hands the offsets to
//     the workers through shared memory, one iteration behind them.
// Shared control state lives in a ring of 3 slots (iteration % 3): a slot is reused only after
// the workers have copied out the tile two iterations back, which needs the control warp to be
// done with it.
constexpr int kRing = 3;

// NW worker warps (+ 1 control warp) per CTA: 12 for short waves (two CTAs per SM: measured best,
// 0.69 vs 0.74 ms for 3 x 8 on C2; 13 and 14 lose to register pressure), 8 when the staging of
// longer waves needs the room
template <int K, int MINB, bool kDelta, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, MINB)
encode_tile_kernel(const EncodeParams p, const uint32_t stage_words, const uint32_t ntiles)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t s_tile[kRing];                   // tile index of the iteration
    __shared__ uint32_t s_mine[kRing][NW];        // words each wave contributes
    __shared__ uint32_t s_cnt[kRing];                    // workers that have reported
    __shared__ uint32_t s_total[kRing];
    __shared__ uint64_t s_off[kRing];                    // tile's exclusive word offset
    __shared__ volatile uint32_t s_flag[kRing];          // = it + 1 once s_off is valid
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool control = warp == NW;

    if (threadIdx.x < kRing) { s_flag[threadIdx.x] = 0; s_cnt[threadIdx.x] = 0; }
    if (threadIdx.x == 0) s_tile[0] = atomicAdd(p.ticket, 1u);
    __syncthreads();

    if (control) {
        for (uint32_t it = 0;; ++it) {
            const int slot = it % kRing;
            // sleeps on a named barrier (one per ring slot) until the tile's last worker arrives
            asm volatile("bar.sync %0, 64;" ::"r"(2 + slot) : "memory");
            const uint32_t tile = s_tile[slot];
            if (tile >= ntiles) break;
            const uint64_t mine = s_total[slot];
            const uint64_t excl = lookback_excl(p.lookback, tile, mine, lane);
            if (lane == 0) {
                s_off[slot] = excl;
                __threadfence_block();
                s_flag[slot] = it + 1;
                if (excl + mine > p.out_cap_words) atomicOr(p.status, kErrCapacity);
                if (tile == ntiles - 1) p.chunk_byte_off[p.nchunks] = (excl + mine) * 4;
            }
            __syncwarp();
        }
        return;
    }

    // ---- workers ---------------------------------------------------------------------------
    uint32_t *const stage0 = smem + (size_t)(2 * warp) * stage_words;   // two staging buffers per warp
    const int16_t *const raw_hi = p.raw + p.raw_samples;
    WaveGeom wg_prev;
    uint32_t nwords_prev = 0;
    bool have_prev = false, ovf_prev = false;
    wg_prev.g = 0xffffffffu;
    for (uint32_t it = 0;; ++it) {
        const int par = it & 1, slot = it % kRing;
        const uint32_t tile = s_tile[slot];
        const bool live = tile < ntiles;
        uint32_t next_ticket = 0;
        WaveGeom wg;
        uint32_t nwords = 0;
        bool have = false, ovf = false;
        if (live) {
            const uint32_t g = tile * NW + warp;
            uint32_t mine = 0;
            if (g < p.nwaves) {
                have = true;
                wg = locate_wave(p, g);
                if (wg.chunk_total) {
                    nwords = (encode_wave<K, false, kDelta>(p.raw + wg.begin, wg.n, raw_hi, lane, stage0 + par * stage_words,
                                                    stage_words, &ovf) + 31u) >> 5;
                    mine = nwords + 1u;
                }
                mine += wg.first;                            // empty chunk: header only
            }
            // next tile: taken as late as possible so that tiles start in ticket order; the
            // atomic's latency hides behind the copy-out below
            if (threadIdx.x == 0) next_ticket = atomicAdd(p.ticket, 1u);
            bool last = false;
            if (lane == 0) {
                s_mine[slot][warp] = mine;
                __threadfence_block();
                if (atomicAdd(&s_cnt[slot], 1u) == NW - 1) {
                    // last worker of the tile: publish the tile's aggregate now
                    __threadfence_block();
                    uint32_t total = 0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) total += s_mine[slot][w];
                    st_relaxed_u64(p.lookback + tile, kFlagAggregate | (uint64_t)total);
                    s_total[slot] = total;
                    s_cnt[slot] = 0;
                    __threadfence_block();
                    last = true;
                }
            }
            if (__shfl_sync(0xffffffffu, last, 0))           // wakes the control warp
                asm volatile("bar.arrive %0, 64;" ::"r"(2 + slot) : "memory");
        } else if (warp == 0) {
            __threadfence_block();
            asm volatile("bar.arrive %0, 64;" ::"r"(2 + slot) : "memory");   // lets the control warp see the end
        }
        // ---- copy out the wave of the previous iteration ------------------------------------
        if (have_prev) {
            const int ps = (it + kRing - 1) % kRing;
            while (s_flag[ps] != it) __nanosleep(40);        // tile offset: normally there long ago
            __threadfence_block();
            const uint32_t v = lane < NW ? s_mine[ps][lane] : 0u;
            const uint32_t loff = __reduce_add_sync(0xffffffffu, lane < warp ? v : 0u);
            const uint64_t off = s_off[ps] + loff;
            const uint32_t rec_words = wg_prev.chunk_total ? nwords_prev + 1u : 0u;
            const bool fits = off + rec_words + wg_prev.first <= p.out_cap_words;
            if (lane == 0 && wg_prev.first) p.chunk_byte_off[wg_prev.chunk] = off * 4;
            if (fits) {
                uint32_t *rec = p.out + off + wg_prev.first;
                if (lane == 0) {
                    if (wg_prev.first) p.out[off] = wg_prev.chunk_total;
                    if (rec_words) rec[0] = nwords_prev;
                }
                if (rec_words) {
                    if (!ovf_prev) {
                        const uint32_t *src = stage0 + (par ^ 1) * stage_words;
                        for (uint32_t i = lane; i < nwords_prev; i += 32) rec[1 + i] = src[i];
                    } else {                                 // larger than the staging: pack in place
                        bool dummy;
                        encode_wave<K, true, kDelta>(p.raw + wg_prev.begin, wg_prev.n, raw_hi, lane, rec + 1, nwords_prev, &dummy);
                    }
                }
            }
            __syncwarp();
        }
        if (!live) break;
        if (threadIdx.x == 0) s_tile[(it + 1) % kRing] = next_ticket;
        wg_prev = wg;
        nwords_prev = nwords;
        have_prev = have;
        ovf_prev = ovf;
        // workers only (the control warp runs on its own clock)
        asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
    }
}

// zig-zag of the 16-bit wrapped difference cur-prev (src/deltaRice.c:57-62, :207-211)
__device__ __forceinline__ uint32_t zigzag_delta(int cur, int prev)
{
    const int t = cur - prev;
    const uint32_t s = (uint32_t)((int)((uint32_t)t << 16) >> 31);
    return (((uint32_t)t << 1) ^ s) & 0xFFFFu;
}

// ======================================================================================
// lane kernel: one LANE per wave (large batches)
// ======================================================================================
// With hundreds of thousands of waves in a batch the parallelism is across waves, as in the
// decoder: every lane encodes its OWN wave sequentially, so nothing is shared inside a wave -
// no per-round warp scan, no stitching of neighbouring lanes' bits, no warp-uniform escape
// handling: delta / zig-zag on packed halves, pair codes, and the multiply-append packer are the
// whole loop (17 instead of 31 instructions per sample).  The price is that a wave's size is only
// known when it is done, so the lane first packs into a worst-case sized SLOT of an HBM scratch
// (32-byte sectors, through a 32-word ring in shared memory), and when the 32 waves of the warp
// task are finished the warp resolves the task's offset by decoupled look-back over TASKS and
// copies the records to their final place, coalesced (the slots are still in L2 for the most part).
//   input:  one 32-byte sector per lane and 16 samples, requested a block ahead;
//   output: ring[word][lane] (bank = lane), flushed as full sectors.
constexpr int kLaneWarps = 17;               // per CTA; two CTAs per SM
constexpr int kLaneRingWords = 32;           // per lane

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
struct Sector { uint32_t w[8]; };
__device__ __forceinline__ Sector ldg_sector(const void *p)
{
    Sector r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_sector(void *p, const uint32_t (&v)[8])
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ uint32_t ldg_cg_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ldg_cg_v4(const uint32_t *p)
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// bit packer of one lane over its whole wave (same multiply-append as Packer); finished words go
// to the lane's ring, full sectors of the ring to the wave's scratch slot
struct LanePacker {
    uint32_t lo, n;             // pending bits in the low n (< 32) bits of lo
    uint32_t wcount;            // words produced so far
    uint32_t flushed;           // words already in the slot (multiple of 8)
    uint32_t ring_b;            // shared address of the lane's ring row 0
    uint32_t *slot;

    __device__ __forceinline__ void put(uint32_t v, uint32_t len)
    {
        const uint64_t a = mul_wide(lo, pow2(len));
        const uint32_t alo = (uint32_t)a | v;
        n += len;
        if (n >= 32u) {
            n -= 32u;
            sts32(ring_b + ((wcount & (kLaneRingWords - 1u)) << 7), __funnelshift_r(alo, (uint32_t)(a >> 32), n));
            ++wcount;
        }
        lo = alo;
    }
    __device__ __forceinline__ void flush_sectors()
    {
        while (wcount - flushed >= 8u) {
            const uint32_t ad = ring_b + ((flushed & (kLaneRingWords - 1u)) << 7);
            uint32_t v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = lds32(ad + 128u * i);
            stg_sector(slot + flushed, v);
            flushed += 8u;
        }
    }
    // closes the wave: the last partial word is left aligned, zero padded (:237-241)
    __device__ __forceinline__ uint32_t finish()
    {
        if (n) {
            uint32_t last;
            asm("shl.b32 %0, %1, %2;" : "=r"(last) : "r"(lo), "r"(32u - n));
            sts32(ring_b + ((wcount & (kLaneRingWords - 1u)) << 7), last);
            ++wcount;
        }
        flush_sectors();
        for (uint32_t i = flushed; i < wcount; ++i) slot[i] = lds32(ring_b + ((i & (kLaneRingWords - 1u)) << 7));
        return wcount;
    }
};

// one sample through the generic code (prologue / epilogue of a wave, escapes)
template <int K>
__device__ __forceinline__ void lane_put_sample(LanePacker &pk, uint32_t u)
{
    uint32_t v, l;
    rice_code<K>(u, v, l);
    pk.put(v, l);
}

// Work is handed out in SLICES: an item = (task of 32 waves, slice of kLaneSliceBlocks * 16 samples),
// tickets run slice-major (every task's slice 0, then every task's slice 1, ...), and a lane's state
// (packer, previous sample, position; the < 8 words still in its ring go to the slot) is parked in
// global memory between slices.  With one long task per warp the SM's warp scheduler favours its
// older CTA: that CTA's tasks finished ~170 us before the other's and half of the machine then ran at
// half its warps and half its IPC (measured with per-task timestamps).  Slices keep all tasks in lock
// step, so every warp has work until the end, whatever the number of tasks per warp.
constexpr uint32_t kLaneSliceBlocks = 32;     // 512 samples per slice
constexpr int      kLaneStateWords  = 6;      // per lane: lo, n, wcount, flushed, previous sample, samples done

template <int K>
__global__ void __launch_bounds__(kLaneWarps * 32, 2)
encode_lane_kernel(const EncodeParams p, uint32_t *const scratch, const uint32_t slot_words, const uint32_t ngroups,
                   const uint32_t nslices, uint32_t *const state, uint32_t *const slice_done, const uint32_t mul_x,
                   const uint32_t neg_prev)
{   // (mul_x, neg_prev): pre-filter mode, delta = (0xFFFF0001, 0xFFFFFFFF), none = (1, 0)
    using C = RiceConst<K>;
    constexpr bool kPairs = C::kPairs;
    constexpr uint32_t M = C::M;
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ring_b = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)warp * (kLaneRingWords * 128u) + (uint32_t)lane * 4u;
    asm volatile("mov.u32 %0, %0;" : "+r"(ring_b));         // keep it in a register (not recomputed per word)
    const int16_t *const raw_hi = p.raw + p.raw_samples;
    const int dmask = (int)neg_prev;                        // -1: delta, 0: none
    const uint32_t nitems = ngroups * nslices;

    while (true) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(p.ticket, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) break;
        const uint32_t slice = item / ngroups;
        const uint32_t grp = item - slice * ngroups;
        const bool last_slice = slice == nslices - 1;
        const uint32_t g = grp * 32u + (uint32_t)lane;
        const bool active = g < p.nwaves;
        WaveGeom wg;
        wg.begin = 0; wg.n = 0; wg.chunk = 0; wg.first = 0; wg.chunk_total = 0; wg.g = g; wg.pad_ = 0;
        if (active) wg = locate_wave(p, g);
        const uint32_t n = wg.chunk_total ? wg.n : 0u;

        LanePacker pk;
        pk.lo = 0; pk.n = 0; pk.wcount = 0; pk.flushed = 0;
        pk.ring_b = ring_b;
        pk.slot = scratch + (size_t)g * slot_words;
        int prevs = 0;                                       // previous sample of the wave
        uint32_t done = 0;                                   // samples of the wave already coded
        uint32_t *const st = state + ((size_t)grp * kLaneStateWords) * 32u + (uint32_t)lane;
        __syncwarp();
        if (slice) {
            // the task's previous slice (handed out ngroups tickets ago) must have parked its state
            if (lane == 0) {
                while (ld_acquire_u32(slice_done + grp) < slice) __nanosleep(100);
            }
            __syncwarp();
            pk.lo = ldg_cg_u32(st);
            pk.n = ldg_cg_u32(st + 32);
            pk.wcount = ldg_cg_u32(st + 64);
            pk.flushed = ldg_cg_u32(st + 96);
            prevs = (int)ldg_cg_u32(st + 128);
            done = ldg_cg_u32(st + 160);
            if (n) for (uint32_t i = pk.flushed; i < pk.wcount; ++i) sts32(ring_b + ((i & (kLaneRingWords - 1u)) << 7), ldg_cg_u32(pk.slot + i));
        }
        const int16_t *q = p.raw + wg.begin + done;
        uint32_t left = n - done;

        // ---- prologue (slice 0): single samples up to the first 32-byte boundary of the input -----
        if (slice == 0) {
            uint32_t pro = (uint32_t)((32u - (uint32_t)(reinterpret_cast<uintptr_t>(q) & 31u)) & 31u) >> 1;
            if (pro > left) pro = left;
            left -= pro;
            done += pro;
            for (; pro; --pro) {
                const int x = *q++;
                lane_put_sample<K>(pk, zigzag_delta(x, prevs & dmask));
                prevs = x;
            }
        }
        // ---- blocks of 16 samples = one sector ----------------------------------------------------
        uint32_t nblk = left >> 4;
        if (nblk > kLaneSliceBlocks) nblk = kLaneSliceBlocks;
        const uint32_t maxblk = __reduce_max_sync(0xffffffffu, nblk);
        uint32_t pw = (uint32_t)prevs << 16;                 // previous sample in the HIGH half
        Sector nxt;
#pragma unroll
        for (int m = 0; m < 8; ++m) nxt.w[m] = 0;
        if (nblk) nxt = ldg_sector(q);
        const uint32_t MMr = opaque(C::MM), NNr = opaque(C::NN);
        for (uint32_t b = 0; b < maxblk; ++b) {
            if (b < nblk) {
                const Sector cur = nxt;
                q += 16;
                if (b + 1 < nblk) nxt = ldg_sector(q);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    // delta + zig-zag on packed halves (src/deltaRice.c:57-62, :207-211)
                    const uint32_t prev = m ? cur.w[m - 1] : pw;
                    const uint32_t X = cur.w[m] * mul_x;
                    const uint32_t Y = mad_lo(prev >> 16, neg_prev, cur.w[m]);
                    const uint32_t D = prmt(Y, X, 0x7610);
                    const uint32_t Sg = prmt(D, 0, 0xbb99);
                    const uint32_t u2 = __vadd2(D, D) ^ Sg;
                    if (kPairs) {
                        if ((u2 & C::HM) == 0u) {            // two samples as one code of <= 30 bits
                            const uint32_t V2 = (u2 | MMr) & NNr;
                            const uint32_t qhi = u2 >> (16 + K), qlo = (u2 >> K) & C::QM;
                            pk.put(mad_lo(V2 & 0xFFFFu, (2u * M) << qhi, V2 >> 16), qlo + qhi + 2u * (K + 1));
                        } else {                             // a quotient >= 8: escape code(s)
                            lane_put_sample<K>(pk, u2 & 0xFFFFu);
                            lane_put_sample<K>(pk, u2 >> 16);
                        }
                    } else {
                        lane_put_sample<K>(pk, u2 & 0xFFFFu);
                        lane_put_sample<K>(pk, u2 >> 16);
                    }
                }
                pw = cur.w[7];
                pk.flush_sectors();
            }
        }
        if (nblk) prevs = (int)pw >> 16;
        done += nblk * 16u;
        left -= nblk * 16u;

        if (!last_slice) {
            // ---- park the lane: pending words to the slot, state to global memory -----------------
            if (n) for (uint32_t i = pk.flushed; i < pk.wcount; ++i) pk.slot[i] = lds32(ring_b + ((i & (kLaneRingWords - 1u)) << 7));
            st[0] = pk.lo;
            st[32] = pk.n;
            st[64] = pk.wcount;
            st[96] = pk.flushed;
            st[128] = (uint32_t)prevs;
            st[160] = done;
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_u32(slice_done + grp, slice + 1u);
            continue;
        }
        // ---- epilogue (last slice): the last < 16 samples ----------------------------------------------
        for (; left; --left) {
            const int x = (q < raw_hi) ? (int)*q : 0;
            ++q;
            lane_put_sample<K>(pk, zigzag_delta(x, prevs & dmask));
            prevs = x;
        }
        const uint32_t nwords = n ? pk.finish() : 0u;
        __syncwarp();

        // ---- the task's 32 records -> their final place --------------------------------------------
        // (the last slices are handed out in task order, so the look-back finds its prefix close by)
        const uint32_t rec_words = (active && wg.chunk_total) ? nwords + 1u : 0u;
        const uint32_t mine = rec_words + (active ? wg.first : 0u);     // empty chunk: header only
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) st_relaxed_u64(p.lookback + grp, kFlagAggregate | (uint64_t)total);
        const uint64_t excl = lookback_excl<400>(p.lookback, grp, total, lane);
        const uint64_t off = excl + incl - mine;
        const bool fits = off + mine <= p.out_cap_words;
        if (!fits && mine) atomicOr(p.status, kErrCapacity);
        if (grp == ngroups - 1 && lane == 31) p.chunk_byte_off[p.nchunks] = (excl + total) * 4;
        if (active && wg.first) p.chunk_byte_off[wg.chunk] = off * 4;
        if (active && fits) {
            if (wg.first) p.out[off] = wg.chunk_total;
            if (rec_words) p.out[off + wg.first] = nwords;
        }
        const uint64_t dst0 = off + wg.first + 1u;
        const uint32_t ncopy = (fits && rec_words) ? nwords : 0u;
        // 16 bytes per lane and load, 512 words of a record per step; the loads of the next record
        // are in flight while the current one is stored (the copy is latency bound otherwise)
        auto load4 = [&](uint4 (&v)[4], int s, uint32_t ns, uint32_t i0) {
            const uint32_t *src = scratch + (size_t)(grp * 32u + (uint32_t)s) * slot_words;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t idx = i0 + (uint32_t)u * 128u + (uint32_t)lane * 4u;
                v[u] = idx < ns ? ldg_cg_v4(src + idx) : make_uint4(0, 0, 0, 0);   // (slots are padded to 8 words)
            }
        };
        auto store4 = [&](const uint4 (&v)[4], uint64_t ds, uint32_t ns, uint32_t i0) {
            uint32_t *dst = p.out + ds;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t idx = i0 + (uint32_t)u * 128u + (uint32_t)lane * 4u;
                if (idx + 4u <= ns) {
                    dst[idx] = v[u].x; dst[idx + 1] = v[u].y; dst[idx + 2] = v[u].z; dst[idx + 3] = v[u].w;
                } else if (idx < ns) {
                    dst[idx] = v[u].x;
                    if (idx + 1 < ns) dst[idx + 1] = v[u].y;
                    if (idx + 2 < ns) dst[idx + 2] = v[u].z;
                }
            }
        };
        uint4 cur4[4], nxt4[4];
        uint32_t ns = __shfl_sync(0xffffffffu, ncopy, 0);
        load4(cur4, 0, ns, 0);
        for (int s2 = 0; s2 < 32; ++s2) {
            const uint64_t ds = __shfl_sync(0xffffffffu, dst0, s2);
            const uint32_t ns_next = s2 < 31 ? __shfl_sync(0xffffffffu, ncopy, (s2 + 1) & 31) : 0u;
            if (s2 < 31) load4(nxt4, s2 + 1, ns_next, 0);
            store4(cur4, ds, ns, 0);
            for (uint32_t i0 = 512u; i0 < ns; i0 += 512u) {     // records of more than 512 words
                load4(cur4, s2, ns, i0);
                store4(cur4, ds, ns, i0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) cur4[u] = nxt4[u];
            ns = ns_next;
        }
        __syncwarp();
    }
}

// ======================================================================================
// segment kernel: one WARP per wave, one LANE per contiguous SEGMENT (the default encoder)
// ======================================================================================
// The codec is bound by the integer pipes, not by HBM: on sm_100 the ALU pipe (LOP3 / SHF / PRMT /
// IADD3) and the FMA pipe (IMAD, VIADD.16x2) each take one warp instruction per two cycles per SM
// sub-partition and IMAD.WIDE / IMAD.HI one per four (tools/ubench_pipes.cu), so the kernel is
// designed around pipe cycles per sample and around keeping many warps resident:
//   * a wave is cut into BLOCKS of 16 samples = one 32-byte sector, counted from the 32-byte aligned
//     address below the wave; a PIECE (<= 32 x nbmax blocks, normally the whole wave) deals its
//     blocks evenly to the lanes as contiguous segments (the first lanes get one block more).  A lane
//     reads its segment sector by sector straight into registers (ld.global.nc.v8, two sectors
//     ahead): no shared-memory staging of the input, full sectors only, and the only irregular
//     blocks of a wave are its first (the samples in front of the wave are zeroed: they then code as
//     K+1 known bits each, which are skipped) and its last (padded with repeats of the last sample:
//     zero deltas, cut off again);
//   * every lane Rice-codes its segment sequentially into a private bit stream in shared memory
//     ([word][lane] rings: conflict free) - no per-round warp scan, no stitching inside the loop.
//     Two samples become one pair code with IMAD.WIDE (the multiply is the shift), appended to a
//     64-bit window with one more IMAD.WIDE; the word under construction is stored unconditionally
//     (a complete word overwrites the partial one), so the loop has no branch;
//   * after the piece ONE warp scan of the lanes' bit counts gives every segment its bit offset in
//     the wave and each lane funnel-shifts its stream into the wave's staging (boundary words are
//     stitched with one shuffle);
//   * tiles of NW waves, the control warp's decoupled look-back and the deferred coalesced copy-out
//     are those of the tile kernel, without its per-iteration barrier among the workers.
// A wave that outgrows its staging (or a lane its ring) is packed straight into its record by
// encode_wave<.., true>.
constexpr int      kSegMaxWorkers = 10;       // worker warps per CTA (+ 1 control warp); two CTAs per SM
constexpr uint32_t kSegMinL       = 1024;     // shorter waves: tile kernel (lanes would idle)
constexpr uint32_t kSegMaxBlocks  = 8;        // blocks per lane and piece

struct SegLaunch {
    uint32_t nworkers;      // worker warps per CTA (+ 1 control warp)
    uint32_t nbmax;         // blocks (of 16 samples) per lane and piece, <= kSegMaxBlocks
    uint32_t lane_words;    // words of one lane's private stream ring: a power of two
    uint32_t stage_words;   // words of a wave's merged stream (nstage buffers per worker)
    uint32_t nstage;        // 2 or 3: a wave is copied out nstage - 1 iterations after it was encoded
    uint32_t ntiles;
    uint32_t *max_words;    // hints for the next call (may be null): [0] largest record, [1] largest lane stream (words)
    // multiplier constants of the inner loop, passed as PARAMETERS: a power of two the compiler can
    // see is strength-reduced to shifts, i.e. moved from the FMA pipe back to the busier ALU pipe
    uint32_t mulx;          // 0xFFFF0001: w * mulx has hi(w) - lo(w) in its high half
    uint32_t p16;           // 2^16
    uint32_t p16mk;         // 2^(16-K)
    uint32_t four;          // 4
    uint32_t dbg;           // DRICE_ENC_SEG_DBG (timing experiments only; output is wrong when set)
    long long *trace;       // DRICE_ENC_SEG_TRACE: clock64 stamps of CTA 0 (diagnostics), else null
};

__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c)
{
    uint64_t d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
// (a ^ b) & c and (a ^ b) & ~c in one LOP3 each; (a & b) | c
__device__ __forceinline__ uint32_t xor_and(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t xor_andn(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x14;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

struct SegPacker {
    uint32_t acc;       // pending bits in the low (b & 31) bits, stale above
    uint32_t b;         // bits produced so far in this piece
    uint32_t base;      // shared address of the lane's ring word 0 (the ring block is aligned to its size)
    uint32_t wmask;     // (lane_words - 1) << 7: byte offset of a ring word
    // appends a code of `len` <= 30 bits given as two disjoint bit fields v0 | v1 (P = 2^len).  The
    // word under construction is stored every time: if the code completes it the store is final,
    // otherwise a later store overwrites it.  (The fields are joined with a three-input OR, not an
    // add, so that they cannot be folded into the multiply as a 64-bit addend.)
    __device__ __forceinline__ void append(uint32_t v0, uint32_t v1, uint32_t P, uint32_t len, const SegLaunch &c)
    {
        const uint64_t a = mad_wide(acc, P, 0ull);
        const uint32_t alo = (uint32_t)a | v0 | v1;          // the low `len` bits of the product are 0
        const uint32_t at = and_or(b * c.four, wmask, base); // word b >> 5 of the ring
        b += len;
        sts32(at, __funnelshift_r(alo, (uint32_t)(a >> 32), b));
        acc = alo;
    }
};

// two samples (both quotients < 8) as one code: VL = packed remainders, QK = packed quotient * M.
template <int K>
__device__ __forceinline__ void seg_put_pair(SegPacker &pk, uint32_t VL, uint32_t QK, const SegLaunch &c)
{
    const uint64_t qs = mad_wide(QK, c.p16mk, 0ull);                         // {q_hi, q_lo << 16}
    const uint32_t q_hi = (uint32_t)(qs >> 32);
    const uint32_t S = ((uint32_t)qs >> 16) + q_hi;                           // q_lo + q_hi (one LEA.HI)
    const uint32_t Ph = __funnelshift_l(0u, (2u << K) << 16, q_hi);           // 2^(16 + len_hi)
    // terminator bits: VIADD.16x2 (FMA pipe); the halves cannot carry (r < M)
    const uint64_t sp = mad_wide(__vadd2(VL, RiceConst<K>::MM), c.p16, 0ull); // {r_hi + M, (r_lo + M) << 16}
    const uint64_t t = mad_wide((uint32_t)sp, Ph, 0ull);                      // high word: (r_lo + M) * 2^len_hi
    const uint32_t P = __funnelshift_l(0u, 1u << (2 * K + 2), S);
    pk.append((uint32_t)(t >> 32), (uint32_t)(sp >> 32), P, S + (2u * K + 2u), c);
}
template <int K>
__device__ __forceinline__ void seg_put_single(SegPacker &pk, uint32_t u, const SegLaunch &c)
{
    uint32_t v, l;
    rice_code<K>(u, v, l);
    pk.append(v, 0u, pow2(l), l, c);
}

// 8 samples = 4 packed words of one lane.  pw = the word before them (previous sample in its high half).
template <int K, bool kDelta>
__device__ __forceinline__ void seg_block(SegPacker &pk, const uint32_t (&w)[4], uint32_t &pw, const SegLaunch &c)
{
    using C = RiceConst<K>;
    constexpr uint32_t LOW = ((1u << K) - 1u) * 0x10001u;
    uint32_t VL[4], QK[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        // delta + zig-zag on packed halves (src/deltaRice.c:57-62, :207-211)
        uint32_t D = w[m];
        if (kDelta) {
            const uint32_t prev = m ? w[m - 1] : pw;
            const uint32_t X = w[m] * c.mulx;                // high half: hi(w) - lo(w)
            const uint32_t Y = w[m] - (prev >> 16);          // low half:  lo(w) - hi(prev)
            D = prmt(Y, X, 0x7610);
        }
        const uint32_t D2 = __vadd2(D, D), Sg = prmt(D, 0, 0xbb99);    // U = D2 ^ Sg
        VL[m] = xor_and(D2, Sg, LOW);
        QK[m] = xor_andn(D2, Sg, LOW);
    }
    pw = w[3];
    if constexpr (C::kPairs) {
        const bool esc = (((QK[0] | QK[1]) | (QK[2] | QK[3])) & C::HM) != 0u;
        if (!__any_sync(__activemask(), esc)) {
#pragma unroll
            for (int m = 0; m < 4; ++m) seg_put_pair<K>(pk, VL[m], QK[m], c);
        } else {                                    // some lane holds a quotient >= 8 (escape code)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (__any_sync(__activemask(), (QK[m] & C::HM) != 0u)) {
                    const uint32_t U = VL[m] | QK[m];
                    seg_put_single<K>(pk, U & 0xFFFFu, c);
                    seg_put_single<K>(pk, U >> 16, c);
                } else {
                    seg_put_pair<K>(pk, VL[m], QK[m], c);
                }
            }
        }
    } else {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const uint32_t U = VL[m] | QK[m];
            seg_put_single<K>(pk, U & 0xFFFFu, c);
            seg_put_single<K>(pk, U >> 16, c);
        }
    }
}

// block B (16 samples from origin + 32 B) of a wave.  Only the wave's first and last block can be ragged
// (samples in front of the wave / behind it inside the 32-byte sector): those two sectors are patched
// once per wave into a 64-byte scratch of the warp in shared memory (seg_patch) and read from there.
__device__ __forceinline__ Sector lds_sector(uint32_t addr)
{
    Sector r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "r"(addr));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "r"(addr));
    return r;
}
__device__ __forceinline__ Sector seg_load(const unsigned char *origin, uint32_t B, uint32_t NB, bool lead, bool tail, uint32_t scratch)
{
    if ((B == 0u && lead) || (B + 1u == NB && tail)) return lds_sector(scratch + (B == 0u && lead ? 0u : 32u));
    return ldg_sector(origin + (size_t)B * 32u);
}
// lanes 0..15: the 16 samples of the wave's first sector, lanes 16..31: of its last sector.  Samples in front
// of the wave become 0, samples behind it repeat the last one (`fill`); nothing outside the wave is read.
__device__ __forceinline__ void seg_patch(const unsigned char *origin, uint32_t NB, uint32_t mis, uint32_t span, uint32_t fill,
                                          uint32_t scratch, int lane)
{
    const uint32_t idx = (lane < 16 ? 0u : (NB - 1u) * 16u) + ((uint32_t)lane & 15u);
    uint32_t v = idx < mis ? 0u : fill;
    if (idx >= mis && idx < span) v = *reinterpret_cast<const uint16_t *>(origin + (size_t)idx * 2u);
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(scratch + (uint32_t)lane * 2u), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// the rare wave that outgrew its staging: out of line, so that it stays out of the instruction cache
template <int K, bool kDelta>
__device__ __noinline__ void encode_wave_in_place(const int16_t *wave, uint32_t n, const int16_t *raw_hi, int lane, uint32_t *dst, uint32_t cap)
{
    bool dummy;
    encode_wave<K, true, kDelta>(wave, n, raw_hi, lane, dst, cap, &dummy);
}

// Shared control state of a CTA lives in a ring of kSegRing slots (iteration % kSegRing).  There is no
// barrier among the workers.  A wave is copied out D = nstage - 1 iterations after it was encoded (the
// merged stream waits in one of the worker's nstage staging buffers), when the control warp has long
// resolved its tile's offset: a worker needs the offset of tile it-D to finish iteration it, and the
// control warp needs every worker's size of that tile, so no worker is ever more than D+1 iterations
// ahead of another and six slots are never reused too early.
// The worker that COMPLETES tile it (the last to report its wave's size) claims the tile of iteration
// it+2: tiles complete in ticket order machine-wide (a tile is claimed two iterations of its CTA's
// slowest worker before it completes), so a look-back finds its predecessors published, and every
// worker knows its next wave one iteration early: it prefetches that wave into L2 while it encodes.
constexpr int kSegRing = 6;

struct SegDeferred {                       // what the copy-out of a wave needs, parked in shared memory
    uint64_t begin;
    uint32_t n, chunk, chunk_total, nwords, flags;   // flags: 1 first, 2 have, 4 overflow
    uint32_t pad_;
};

template <int K, bool kDelta>
__global__ void __launch_bounds__((kSegMaxWorkers + 1) * 32, 2)
encode_seg_kernel(const EncodeParams p, const SegLaunch sl)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t s_tile[8];                        // tile of iteration it (ring of 8)
    __shared__ volatile uint32_t s_tag[8];                // = it + 1 once s_tile[it & 7] is valid
    __shared__ uint32_t s_mine[kSegRing][kSegMaxWorkers]; // words each wave contributes
    __shared__ uint32_t s_cnt[kSegRing];
    __shared__ uint32_t s_total[kSegRing];
    __shared__ uint64_t s_off[kSegRing];
    __shared__ volatile uint32_t s_flag[kSegRing];        // = it + 1 once s_off is valid
    __shared__ SegDeferred s_def[kSegMaxWorkers][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NW = (int)sl.nworkers;
    const bool control = warp == NW;
    const uint32_t ntiles = sl.ntiles;
    const uint32_t D = sl.nstage - 1u;                    // iterations between encoding a wave and copying it out

    if (threadIdx.x < kSegRing) { s_flag[threadIdx.x] = 0; s_cnt[threadIdx.x] = 0; }
    if (threadIdx.x < 8) s_tag[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_tile[0] = atomicAdd(p.ticket, 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        s_tile[1] = atomicAdd(p.ticket, 1u);              // (after the machine's first round of claims, roughly)
        s_tag[0] = 1;
        s_tag[1] = 2;
    }
    __syncthreads();

    if (control) {
        for (uint32_t it = 0;; ++it) {
            const int slot = it % kSegRing;
            // sleeps on a named barrier (one per ring slot) until the tile's last worker arrives
            asm volatile("bar.sync %0, 64;" ::"r"(2 + slot) : "memory");
            const uint32_t tile = s_tile[it & 7];
            if (tile >= ntiles) break;
            const uint64_t mine = s_total[slot];
            const uint64_t excl = lookback_excl(p.lookback, tile, mine, lane);
            if (lane == 0) {
                s_off[slot] = excl;
                __threadfence_block();
                s_flag[slot] = it + 1;
                if (excl + mine > p.out_cap_words) atomicOr(p.status, kErrCapacity);
                if (tile == ntiles - 1) p.chunk_byte_off[p.nchunks] = (excl + mine) * 4;
            }
            __syncwarp();
        }
        return;
    }

    // ---- workers ---------------------------------------------------------------------------
    const uint32_t stage_words = sl.stage_words;
    const uint32_t ring_bytes = sl.lane_words * 128u;                          // one warp's 32 lane rings
    const uint32_t dyn = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t rings0 = (dyn + ring_bytes - 1u) & ~(ring_bytes - 1u);      // ring blocks are aligned to their size
    const uint32_t stage_s0 = rings0 + (uint32_t)NW * ring_bytes + (uint32_t)warp * (sl.nstage * stage_words * 4u + 64u);
    const uint32_t scratch = stage_s0 + sl.nstage * stage_words * 4u;         // the wave's ragged first / last sector
    const int16_t *const raw_hi = p.raw + p.raw_samples;
    SegPacker pk;
    pk.base = rings0 + (uint32_t)warp * ring_bytes + (uint32_t)lane * 4u;
    pk.wmask = (sl.lane_words - 1u) << 7;
    const uint32_t ring_bits = (sl.lane_words - 1u) * 32u;                     // a lane past this has overflowed

    uint32_t maxw = 0, maxlane = 0, novf = 0;
    uint32_t buf = 0;                                    // staging buffer of this iteration: it % nstage
    for (uint32_t it = 0;; ++it) {
        const int slot = it % kSegRing;
        const bool tr = sl.trace && blockIdx.x == 0 && it >= 8u && it < 24u;
        long long *const trp = sl.trace + ((size_t)warp * 16u + (it - 8u)) * 8u;
#define SEG_STAMP(k) do { if (tr && lane == 0) trp[k] = clock64(); } while (0)
        SEG_STAMP(0);
        while (s_tag[it & 7] != it + 1) __nanosleep(20);     // (claimed when tile it-2 completed: long ago)
        __threadfence_block();
        const uint32_t tile = s_tile[it & 7];
        const bool live = tile < ntiles;
        SegDeferred df;
        df.begin = 0; df.n = 0; df.chunk = 0; df.chunk_total = 0; df.nwords = 0; df.flags = 0; df.pad_ = 0;
        if (live) {
            const uint32_t g = tile * (uint32_t)NW + (uint32_t)warp;
            uint32_t mine = 0;
            if (g < p.nwaves) {
                const WaveGeom wg = locate_wave(p, g);

                df.begin = wg.begin; df.n = wg.n; df.chunk = wg.chunk; df.chunk_total = wg.chunk_total;
                df.flags = 2u | wg.first;
                bool ovf = false;
                uint32_t nwords = 0;
                if (wg.chunk_total) {
                    const uintptr_t wa = reinterpret_cast<uintptr_t>(p.raw + wg.begin);
                    const unsigned char *const origin = reinterpret_cast<const unsigned char *>(wa & ~(uintptr_t)31);
                    const uint32_t mis = (uint32_t)(wa & 31u) >> 1;            // samples between origin and the wave
                    const uint32_t span = mis + wg.n;
                    const uint32_t NB = (span + 15u) >> 4;                     // blocks of the wave
                    SEG_STAMP(1);
                    const bool lead = mis != 0u, tail = (span & 15u) != 0u;
                    if (lead || tail) {
                        uint32_t fill = 0;
                        if (kDelta && tail) fill = (uint32_t)(uint16_t)p.raw[wg.begin + wg.n - 1u];
                        seg_patch(origin, NB, mis, span, fill, scratch, lane);
                        __syncwarp();
                    }
                    const uint32_t stage_s = stage_s0 + buf * stage_words * 4u;
                    uint32_t Bw = 0;                 // bits of the wave merged so far
                    uint32_t carry = 0;              // the wave's last, still partial word (left aligned)
                    for (uint32_t blk0 = 0; blk0 < NB; blk0 += 32u * sl.nbmax) {
                        // ---- this lane's segment of the piece: blocks [start, start + mb) ------------------
                        const uint32_t pnb = NB - blk0 < 32u * sl.nbmax ? NB - blk0 : 32u * sl.nbmax;
                        const uint32_t q = pnb >> 5, r = pnb & 31u;
                        const uint32_t mb = q + ((uint32_t)lane < r ? 1u : 0u);
                        const uint32_t start = blk0 + q * (uint32_t)lane + ((uint32_t)lane < r ? (uint32_t)lane : r);
                        const uint32_t nact = q ? 32u : r;                        // lanes that hold blocks
                        pk.acc = 0;
                        pk.b = 0;
                        if (mb) {
                            uint32_t pw = 0;
                            if (kDelta && start) pw = (uint32_t)*reinterpret_cast<const uint16_t *>(origin + (size_t)start * 32u - 2u) << 16;
                            Sector c0 = seg_load(origin, start, NB, lead, tail, scratch), c1 = c0, c2 = c0;
                            if (mb > 1u) c1 = seg_load(origin, start + 1u, NB, lead, tail, scratch);
                            if (mb > 2u) c2 = seg_load(origin, start + 2u, NB, lead, tail, scratch);
                            if (tr && lane == 0) { trp[2] = clock64(); trp[7] = (long long)(c0.w[0] & 1u) + clock64(); }
                            for (uint32_t bl = 0; bl < mb; ++bl) {
                                Sector cur = c0;
                                c0 = c1;
                                c1 = c2;
                                if (bl + 3u < mb) c2 = seg_load(origin, start + bl + 3u, NB, lead, tail, scratch);
#pragma unroll 1
                                for (int half = 0; half < 2; ++half) {           // (one copy of the code: it has to stay in the instruction cache)
                                    const uint32_t w4[4] = {cur.w[0], cur.w[1], cur.w[2], cur.w[3]};
                                    seg_block<K, kDelta>(pk, w4, pw, sl);
                                    cur.w[0] = cur.w[4]; cur.w[1] = cur.w[5]; cur.w[2] = cur.w[6]; cur.w[3] = cur.w[7];
                                }
                            }
                            if (pk.b & 31u) {
                                uint32_t last;
                                asm("shl.b32 %0, %1, %2;" : "=r"(last) : "r"(pk.acc), "r"(32u - (pk.b & 31u)));
                                sts32(and_or(pk.b << 2, pk.wmask, pk.base), last);
                            }
                        }
                        __syncwarp();
                        SEG_STAMP(3);
                        maxlane = pk.b > maxlane ? pk.b : maxlane;
                        // ---- bit offsets of the segments in the wave ---------------------------------------
                        const uint32_t skip = (start == 0u && mb) ? mis * (K + 1u) : 0u;    // the zeroed samples in front
                        uint32_t nbits = 0;
                        if (mb) {
                            nbits = pk.b - skip;
                            if (start + mb == NB) nbits -= (NB * 16u - span) * (K + 1u);  // the repeats behind the wave
                        }
                        uint32_t inc = nbits;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                            if (lane >= d) inc += t;
                        }
                        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
                        const uint32_t dbit = Bw + inc - nbits;                  // first bit of this lane's segment
                        // every segment but the last holds >= 32 bits (a word then has at most two contributors)
                        // and no lane has run around its ring
                        const bool shape_ok = __all_sync(0xffffffffu, pk.b <= ring_bits && (!mb || (uint32_t)lane + 1u >= nact || nbits >= 32u));
                        if (!shape_ok || ((Bw + total + 31u) >> 5) + 1u > stage_words) ovf = true;
                        if (!ovf && !(sl.dbg & 1u)) {
                            // ---- merge: funnel-shift the lane's stream to its place ------------------------
                            uint32_t out_frag = 0, head = 0, last = 0;
                            const uint32_t ebit = dbit + nbits;
                            const uint32_t j0 = dbit >> 5, j1 = (ebit - 1u) >> 5;
                            const bool tail_partial = (ebit & 31u) != 0u;
                            const bool single = j0 == j1;
                            const int32_t t = (int32_t)skip - (int32_t)dbit;
                            const uint32_t tr = (uint32_t)t & 31u;
                            const int32_t m0 = (int32_t)j0 + (t >> 5);           // stream word under stage word j0
                            uint32_t pa = pk.base + (uint32_t)(m0 + 1) * 128u;    // address of stream word m0 + 1
                            uint32_t sa = stage_s + j0 * 4u;
                            if (nbits) {
                                const uint32_t tmask = tail_partial ? ~(0xFFFFFFFFu >> (ebit & 31u)) : 0xFFFFFFFFu;
                                uint32_t cur = m0 >= 0 ? lds32(pa - 128u) : 0u;
                                uint32_t nx = lds32(pa);
                                head = __funnelshift_l(nx, cur, tr) & (0xFFFFFFFFu >> (dbit & 31u));
                                if (single) head &= tmask;
                                uint32_t left = j1 - j0;                          // words after the first
                                while (left > 4u) {                               // four words per step
                                    const uint32_t a1 = lds32(pa + 128u), a2 = lds32(pa + 256u), a3 = lds32(pa + 384u), a4 = lds32(pa + 512u);
                                    sts32(sa + 4u, __funnelshift_l(a1, nx, tr));
                                    sts32(sa + 8u, __funnelshift_l(a2, a1, tr));
                                    sts32(sa + 12u, __funnelshift_l(a3, a2, tr));
                                    sts32(sa + 16u, __funnelshift_l(a4, a3, tr));
                                    nx = a4;
                                    pa += 512u;
                                    sa += 16u;
                                    left -= 4u;
                                }
                                while (left) {
                                    cur = nx;
                                    pa += 128u;
                                    sa += 4u;
                                    nx = lds32(pa);
                                    const uint32_t f = __funnelshift_l(nx, cur, tr);
                                    if (--left) sts32(sa, f); else last = f & tmask;
                                }
                                if (tail_partial && !single) out_frag = last;
                            }
                            uint32_t in_frag = __shfl_up_sync(0xffffffffu, out_frag, 1);
                            if (lane == 0) in_frag = carry;
                            if (nbits) {
                                const uint32_t w0 = head | in_frag;
                                if (single) {
                                    if (tail_partial) out_frag = w0; else sts32(stage_s + j0 * 4u, w0);
                                } else {
                                    sts32(stage_s + j0 * 4u, w0);
                                    if (!tail_partial) sts32(stage_s + j1 * 4u, last);
                                }
                            }
                            carry = __shfl_sync(0xffffffffu, out_frag, (int)nact - 1);
                        }
                        Bw += total;
                        __syncwarp();
                    }
                    SEG_STAMP(4);
                    if (!ovf && (Bw & 31u) && lane == 0) sts32(stage_s + (Bw >> 5) * 4u, carry);
                    nwords = (Bw + 31u) >> 5;
                    mine = nwords + 1u;
                    maxw = nwords > maxw ? nwords : maxw;
                    __syncwarp();
                }
                df.nwords = nwords;
                if (ovf) { df.flags |= 4u; ++novf; }
                mine += wg.first;                            // empty chunk: header only
            }
            bool last = false;
            if (lane == 0) {
                s_mine[slot][warp] = mine;
                __threadfence_block();
                if (atomicAdd(&s_cnt[slot], 1u) == (uint32_t)NW - 1u) {
                    // last worker of the tile: publish the tile's aggregate, claim the tile of iteration it + 2
                    __threadfence_block();
                    uint32_t total = 0;
                    for (int w = 0; w < NW; ++w) total += s_mine[slot][w];
                    st_relaxed_u64(p.lookback + tile, kFlagAggregate | (uint64_t)total);
                    s_total[slot] = total;
                    s_cnt[slot] = 0;
                    s_tile[(it + 2) & 7] = atomicAdd(p.ticket, 1u);
                    __threadfence_block();
                    s_tag[(it + 2) & 7] = it + 3;
                    last = true;
                }
            }
            if (__shfl_sync(0xffffffffu, last, 0))           // wakes the control warp
                asm volatile("bar.arrive %0, 64;" ::"r"(2 + slot) : "memory");
        } else if (warp == 0) {
            __threadfence_block();
            asm volatile("bar.arrive %0, 64;" ::"r"(2 + slot) : "memory");   // lets the control warp see the end
        }
        if (lane == 0) s_def[warp][buf] = df;
        __syncwarp();
        SEG_STAMP(5);
        // ---- copy out the wave(s) encoded D iterations ago (all that are left once the tiles have run out) ----
        for (uint32_t back = D; back >= (live ? D : 1u); --back) {
            if (it < back) continue;
            const uint32_t jt = it - back;                       // iteration whose wave goes out
            const uint32_t jb = jt % sl.nstage;
            const SegDeferred d = s_def[warp][jb];
            if (!(d.flags & 2u)) continue;
            const int psl = jt % kSegRing;
            while (s_flag[psl] != jt + 1) __nanosleep(400);      // tile offset: normally there long ago
            __threadfence_block();
            const uint32_t v = lane < NW ? s_mine[psl][lane] : 0u;
            const uint32_t loff = __reduce_add_sync(0xffffffffu, lane < warp ? v : 0u);
            const uint64_t off = s_off[psl] + loff;
            const uint32_t first = d.flags & 1u;
            const uint32_t rec_words = d.chunk_total ? d.nwords + 1u : 0u;
            const bool fits = off + rec_words + first <= p.out_cap_words;
            if (lane == 0 && first) p.chunk_byte_off[d.chunk] = off * 4;
            if (fits) {
                uint32_t *rec = p.out + off + first;
                if (lane == 0) {
                    if (first) p.out[off] = d.chunk_total;
                    if (rec_words) rec[0] = d.nwords;
                }
                if (rec_words) {
                    if (sl.dbg & 2u) {
                    } else if (!(d.flags & 4u)) {
                        uint32_t sa = stage_s0 + jb * stage_words * 4u + (uint32_t)lane * 4u;
                        uint32_t *dst = rec + 1 + lane;
                        uint32_t i = lane;
                        for (; i + 96u < d.nwords; i += 128u, sa += 512u, dst += 128) {
                            const uint32_t a0 = lds32(sa), a1 = lds32(sa + 128u), a2 = lds32(sa + 256u), a3 = lds32(sa + 384u);
                            dst[0] = a0; dst[32] = a1; dst[64] = a2; dst[96] = a3;
                        }
                        for (; i < d.nwords; i += 32u, sa += 128u, dst += 32) *dst = lds32(sa);
                    } else {                                 // larger than the staging: pack in place
                        encode_wave_in_place<K, kDelta>(p.raw + d.begin, d.n, raw_hi, lane, rec + 1, d.nwords);
                    }
                }
            }
            __syncwarp();
            if (back == 1u) break;
        }
        SEG_STAMP(6);
        if (!live) break;
        buf = buf + 1u == sl.nstage ? 0u : buf + 1u;
    }
    maxlane = __reduce_max_sync(0xffffffffu, maxlane);
    if (sl.max_words && lane == 0 && maxw) {
        atomicMax(sl.max_words, maxw);
        atomicMax(sl.max_words + 1, (maxlane + 31u) >> 5);
        if (novf) atomicAdd(sl.max_words + 2, novf);      // waves that took the in-place path (diagnostics)
    }
}

// ======================================================================================
// multi-tile kernel (waves longer than one tile): generic per-sample code path
// ======================================================================================
__device__ __forceinline__ int4 ld_stream_v4(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <int K>
__device__ __forceinline__ uint32_t rice_code_packed(uint32_t u)
{
    uint32_t v, l;
    rice_code<K>(u, v, l);
    return v | (l << 24);
}

// Codes of the 16-sample slot whose first sample has wave-relative index s0 (may be negative
// or run past n: such samples get length 0).  Returns the slot's bit total.  Slots are
// aligned to memory (32 bytes), not to the wave.
template <int K>
__device__ __forceinline__ uint32_t slot_codes(const int16_t *wave, int64_t s0, uint32_t n,
                                               uint32_t (&cv)[kSamplesPerThread], int dmask)
{   // dmask: -1 = delta pre-filter, 0 = none (the samples are coded as they are)
    const int64_t lo64 = -s0, hi64 = (int64_t)n - s0;
    const int jlo = lo64 > 0 ? (int)(lo64 < S ? lo64 : S) : 0;
    const int jhi = hi64 < S ? (int)(hi64 > 0 ? hi64 : 0) : S;
    uint32_t T = 0;
    if (jhi <= jlo) {
#pragma unroll
        for (int j = 0; j < S; ++j) cv[j] = 0;
        return 0;
    }
    const int16_t *sp = wave + s0;
    int x[S + 1];
    if (jlo == 0 && jhi == S) {
        const int4 *vp = reinterpret_cast<const int4 *>(sp);
        const int4 a = ld_stream_v4(vp), b = ld_stream_v4(vp + 1);
        const int wd[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        x[0] = (s0 > 0) ? (int)sp[-1] : 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            x[2 * m + 1] = (int)(short)(wd[m] & 0xFFFF);
            x[2 * m + 2] = wd[m] >> 16;
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
            cv[j] = rice_code_packed<K>(zigzag_delta(x[j + 1], x[j] & dmask));
            T += cv[j] >> 24;
        }
    } else {
        x[0] = (jlo == 0 && s0 > 0) ? (int)sp[-1] : 0;
#pragma unroll
        for (int j = 0; j < S; ++j) x[j + 1] = (j >= jlo && j < jhi) ? (int)sp[j] : 0;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const uint32_t c = rice_code_packed<K>(zigzag_delta(x[j + 1], x[j] & dmask));
            cv[j] = (j >= jlo && j < jhi) ? c : 0u;
            T += cv[j] >> 24;
        }
    }
    return T;
}

// exclusive scan of v over the block; *total receives the block sum.  swarp: >= 33 words.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *swarp, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) swarp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t ws = lane < nwarps ? swarp[lane] : 0u;
        uint32_t wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        swarp[lane] = wi - ws;              // exclusive warp offsets
        if (lane == 31) swarp[32] = wi;     // block total
    }
    __syncthreads();
    *total = swarp[32];
    return swarp[warp] + inc - v;
}

// Packs the slot's codes into shared words starting at tile-local bit position b0.  Words are
// written by the thread that STARTS them; the leading fragment a thread contributes to a word
// started by a predecessor goes to a side array and is OR-ed in by that word's owner.
__device__ __forceinline__ void pack_slot(const uint32_t (&cv)[kSamplesPerThread], uint32_t b0,
                                          uint32_t T, uint32_t *sbits, uint32_t *shead,
                                          bool &owner, uint32_t &tail, uint32_t &wt)
{
    const uint32_t fill0 = b0 & 31u;
    const uint32_t w0 = b0 >> 5;
    uint32_t fill = fill0, acc = 0;
    uint32_t *dst = fill0 ? (shead + threadIdx.x) : (sbits + w0);
    uint32_t *nxt = sbits + w0 + 1;
#pragma unroll
    for (int j = 0; j < kSamplesPerThread; ++j) {
        const uint32_t len = cv[j] >> 24;
        const uint32_t V = len ? ((cv[j] & 0xFFFFFFu) << (32u - len)) : 0u;   // left aligned
        acc |= V >> fill;
        const uint32_t spill = __funnelshift_r(0u, V, fill);                  // V << (32-fill), 0 if fill==0
        fill += len;
        if (fill >= 32u) {
            *dst = acc;
            dst = nxt;
            ++nxt;
            acc = spill;
            fill -= 32u;
        }
    }
    const uint32_t end = b0 + T;
    const bool crossed = (end >> 5) > w0;
    owner = (T > 0u) && ((end & 31u) != 0u) && (crossed || fill0 == 0u);
    tail = acc;
    wt = end >> 5;
    if (T > 0u && !crossed && fill0 != 0u) shead[threadIdx.x] = acc;   // lies inside a foreign word
}

// packing sweep of one wave by a whole CTA: codes are recomputed tile by tile, packed in
// shared memory and completed words streamed to rec[1..]
template <int K>
__device__ __forceinline__ void pack_wave_streaming(const int16_t *wave, uint32_t n, uint32_t *rec, uint32_t *smem, int dmask)
{
    const int NT = blockDim.x;
    const int tid = threadIdx.x;
    uint32_t *sbits = smem;                  // NT*13 + 1 (worst case 12.5 words per slot)
    uint32_t *shead = sbits + NT * 13 + 1;   // NT
    uint32_t *sboff = shead + NT;            // NT
    uint32_t *swarp = sboff + NT;            // 33
    __shared__ uint32_t s_carry;
    const int mis = (int)((reinterpret_cast<uintptr_t>(wave) & 31u) >> 1);
    const uint32_t nslots = (uint32_t)(((uint64_t)mis + n + S - 1) / S);
    const uint32_t ntiles = (nslots + NT - 1) / NT;
    uint32_t cv[S];
    if (tid == 0) s_carry = 0;
    __syncthreads();
    uint64_t P = 0;                            // bits emitted so far
    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint32_t T = slot_codes<K>(wave, ((int64_t)t * NT + tid) * S - mis, n, cv, dmask);
        uint32_t total;
        const uint32_t boff = block_excl_scan(T, swarp, &total);
        const uint32_t pre = (uint32_t)(P & 31u);          // bits already in word 0 (carry)
        const uint32_t b0 = pre + boff;
        bool owner;
        uint32_t tail, wt;
        shead[tid] = 0;
        sboff[tid] = b0;
        pack_slot(cv, b0, T, sbits, shead, owner, tail, wt);
        __syncthreads();
        if (owner) {
            for (uint32_t j = tid + 1; j < (uint32_t)NT && (sboff[j] >> 5) == wt; ++j) tail |= shead[j];
            sbits[wt] = tail;
        }
        if (tid == 0 && pre) {                 // word 0 was started by the previous tile
            uint32_t c = s_carry;
            for (uint32_t j = 0; j < (uint32_t)NT && (sboff[j] >> 5) == 0; ++j) c |= shead[j];
            sbits[0] = c;
        }
        __syncthreads();
        const uint32_t bend = pre + total;
        const uint32_t full = bend >> 5;
        uint32_t *dstw = rec + 1 + (P >> 5);
        for (uint32_t w = tid; w < full; w += NT) dstw[w] = sbits[w];
        if (tid == 0) s_carry = (bend & 31u) ? sbits[full] : 0u;
        P += total;
        __syncthreads();
    }
    if (tid == 0 && (P & 31u)) rec[1 + (P >> 5)] = s_carry;
}

// waves longer than one warp-kernel wave: one CTA per wave, sizing sweep + look-back + packing sweep
template <int K>
__global__ void __launch_bounds__(kEncMaxThreads)
encode_multi_kernel(const EncodeParams p, const int dmask)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int NT = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ uint32_t s_ticket;
    __shared__ uint64_t s_excl;

    if (tid == 0) s_ticket = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const uint32_t g = s_ticket;
    if (g >= p.nwaves) return;
    const WaveGeom wg = locate_wave(p, g);
    const int16_t *wave = p.raw + wg.begin;
    const int mis = (int)((reinterpret_cast<uintptr_t>(wave) & 31u) >> 1);
    const uint32_t nslots = (uint32_t)(((uint64_t)mis + wg.n + S - 1) / S);
    const uint32_t ntiles = (nslots + NT - 1) / NT;

    uint32_t cv[S];
    // ---- sizing sweep ----------------------------------------------------------------
    uint64_t bits = 0;
    for (uint32_t t = 0; t < ntiles; ++t)
        bits += slot_codes<K>(wave, ((int64_t)t * NT + tid) * S - mis, wg.n, cv, dmask);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, d);
    uint64_t *s64 = reinterpret_cast<uint64_t *>(smem);
    if (lane == 0) s64[warp] = bits;
    __syncthreads();
    uint64_t tot = 0;
    for (int w = 0; w < (NT >> 5); ++w) tot += s64[w];
    __syncthreads();
    const uint32_t nwords = (uint32_t)((tot + 31u) >> 5);
    const uint32_t rec_words = wg.chunk_total ? nwords + 1u : 0u;
    const uint64_t mine = (uint64_t)rec_words + wg.first;
    if (tid == 0) st_relaxed_u64(p.lookback + g, (g == 0 ? kFlagPrefix : kFlagAggregate) | mine);

    if (warp == 0) {
        const uint64_t excl = lookback_excl(p.lookback, g, mine, lane);
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const uint64_t off = s_excl;
    const bool fits = off + mine <= p.out_cap_words;
    if (tid == 0) {
        if (!fits) atomicOr(p.status, kErrCapacity);
        if (wg.first) p.chunk_byte_off[wg.chunk] = off * 4;
        if (g == p.nwaves - 1) p.chunk_byte_off[p.nchunks] = (off + mine) * 4;
    }
    if (!fits) return;
    uint32_t *rec = p.out + off + wg.first;
    if (tid == 0) {
        if (wg.first) p.out[off] = wg.chunk_total;
        if (rec_words) rec[0] = nwords;
    }
    if (rec_words == 0) return;
    pack_wave_streaming<K>(wave, wg.n, rec, smem, dmask);
}

template <int K>
int launch_k(const EncodeParams &p, const EncodeMode &md, uint32_t max_wave_len, cudaStream_t st)
{
    const int g_num_sms = device_sm_count();
    const size_t smem_multi = (size_t)(kEncMaxThreads * 13 + 1 + 2 * kEncMaxThreads + 33 + 3) * sizeof(uint32_t);
    if (max_wave_len > (uint32_t)kEncTileMaxL) {
        encode_multi_kernel<K><<<p.nwaves, kEncMaxThreads, smem_multi, st>>>(p, md.delta ? -1 : 0);
        return 1;
    }
    // large batches: one lane per wave (needs a worst-case sized scratch slot per wave)
    if (md.lane_scratch && md.lane_slot_words) {
        static DeviceOnce attr_lane;
        const size_t smem_lane = (size_t)kLaneWarps * kLaneRingWords * 128;
        if (attr_lane.first()) {
            cudaFuncSetAttribute(encode_lane_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lane);
            cudaFuncSetAttribute(encode_lane_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        }
        const uint32_t ngroups = (p.nwaves + 31u) / 32u;
        uint32_t nslices = (max_wave_len / 16u + kLaneSliceBlocks - 1u) / kLaneSliceBlocks;
        if (nslices < 1u) nslices = 1u;
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, encode_lane_kernel<K>, kLaneWarps * 32, smem_lane);
        if (occ < 1) occ = 1;
        if (occ > 2) occ = 2;
        uint32_t grid = (uint32_t)(occ * g_num_sms);
        const uint32_t need = (ngroups + kLaneWarps - 1) / kLaneWarps;
        if (grid > need) grid = need;
        encode_lane_kernel<K><<<grid, kLaneWarps * 32, smem_lane, st>>>(p, md.lane_scratch, md.lane_slot_words, ngroups, nslices,
                                                                      md.lane_state, md.lane_slice_done,
                                                                      md.delta ? 0xFFFF0001u : 1u, md.delta ? 0xFFFFFFFFu : 0u);
        return 1;
    }
    // the segment kernel: waves of kSegMinL .. kEncTileMaxL samples
    {
        static const bool seg_on = [] { const char *e = getenv("DRICE_ENC_SEG"); return e && atoi(e) != 0; }();
        if (seg_on && max_wave_len >= kSegMinL) {
            SegLaunch sl{};
            // blocks of 16 samples per lane and piece: the whole wave in one piece when that needs <= kSegMaxBlocks
            const uint32_t nb_wave = (max_wave_len + 15u + 15u) / 16u;                 // (+15: misalignment)
            const uint32_t np = (nb_wave + 32u * kSegMaxBlocks - 1u) / (32u * kSegMaxBlocks);
            sl.nbmax = ((nb_wave + np - 1u) / np + 31u) / 32u;
            const uint32_t seg = sl.nbmax * 16u;
            const uint32_t worst = (25u * max_wave_len + 31u) / 32u + 8u;
            // words per record: the previous batch's largest (+ margin), else room for 10 bits per sample
            uint32_t rec = md.seg_words_hint ? md.seg_words_hint + md.seg_words_hint / 16u + 16u
                                             : (10u * max_wave_len + 31u) / 32u + 16u;
            static const long stage_env = [] { const char *e = getenv("DRICE_ENC_STAGE_WORDS"); return e ? atol(e) : 0l; }();
            if (stage_env > 0) rec = (uint32_t)stage_env;
            if (rec > worst) rec = worst;
            if (rec < 64u) rec = 64u;
            sl.stage_words = (rec + 3u) & ~3u;
            // a lane's ring: the previous batch's longest lane stream (+ margin), else room for 12 bits per
            // sample; a power of two, at most the worst case of a segment
            uint32_t lw = 16u;
            const uint32_t lane_need = md.seg_lane_hint ? md.seg_lane_hint + md.seg_lane_hint / 8u + 3u : (12u * seg) / 32u + 3u;
            const uint32_t lane_worst = (25u * seg + 31u) / 32u + 2u;
            while (lw < lane_need && lw < lane_worst) lw <<= 1;
            sl.lane_words = lw;
            static const long nstage_env = [] { const char *e = getenv("DRICE_ENC_SEG_STAGES"); return e ? atol(e) : 0l; }();
            sl.nstage = nstage_env == 2 ? 2u : (nstage_env == 3 ? 3u : (sl.stage_words <= 1024u ? 3u : 2u));
            const size_t warp_bytes = (size_t)lw * 128u + (size_t)sl.nstage * sl.stage_words * 4u + 64u;
            const size_t avail = (227u * 1024u) / 2u - 1024u - (size_t)lw * 128u;      // two CTAs per SM; static + alignment slack
            static const long nw_env = [] { const char *e = getenv("DRICE_ENC_SEG_WARPS"); return e ? atol(e) : 0l; }();
            uint32_t nw = (uint32_t)(avail / warp_bytes);
            if (nw > (uint32_t)kSegMaxWorkers) nw = kSegMaxWorkers;
            if (nw_env > 0 && (uint32_t)nw_env < nw) nw = (uint32_t)nw_env;
            if (nw >= 2u) {
                if (nw > p.nwaves) nw = p.nwaves;
                sl.nworkers = nw;
                sl.ntiles = (p.nwaves + nw - 1u) / nw;
                sl.max_words = md.seg_max_words;
                sl.mulx = 0xFFFF0001u;
                sl.p16 = 65536u;
                sl.p16mk = 65536u >> (K <= 7 ? K : 0);
                sl.four = 4u;
                static const long dbg_env = [] { const char *e = getenv("DRICE_ENC_SEG_DBG"); return e ? atol(e) : 0l; }();
                sl.dbg = (uint32_t)dbg_env;
                static const bool debug = getenv("DRICE_DEBUG") != nullptr;
                if (debug)
                    fprintf(stderr, "[drice] encode_seg K=%d L=%u waves=%u: %u workers/CTA, nbmax=%u, lane ring %u words, staging %u x %u words, "
                                    "hints rec=%u lane=%u, smem %zu B\n", K, max_wave_len, p.nwaves, nw, sl.nbmax, lw, sl.nstage, sl.stage_words,
                            md.seg_words_hint, md.seg_lane_hint, warp_bytes * nw + (size_t)lw * 128u);
                const size_t smem = warp_bytes * nw + (size_t)lw * 128u;
                auto launch_seg = [&](auto kernel, DeviceOnce &once) {
                    if (once.first()) {
                        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
                        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                    }
                    int occ = 0;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, (int)(nw + 1u) * 32, smem);
                    if (occ < 1) occ = 1;
                    uint32_t grid = (uint32_t)(device_sm_count() * occ);
                    if (grid > sl.ntiles) grid = sl.ntiles;
                    kernel<<<grid, (nw + 1u) * 32u, smem, st>>>(p, sl);
                };
                static const char *trace_path = getenv("DRICE_ENC_SEG_TRACE");
                static long long *d_trace = nullptr;
                const size_t trace_n = (size_t)kSegMaxWorkers * 16u * 8u;
                if (trace_path && !d_trace) cudaMalloc((void **)&d_trace, trace_n * sizeof(long long));
                if (trace_path && d_trace) cudaMemsetAsync(d_trace, 0, trace_n * sizeof(long long), st);
                sl.trace = trace_path ? d_trace : nullptr;
                static DeviceOnce once_d, once_n;
                if (md.delta) launch_seg(encode_seg_kernel<K, true>, once_d);
                else launch_seg(encode_seg_kernel<K, false>, once_n);
                if (trace_path && d_trace) {                 // diagnostics only: synchronous
                    std::vector<long long> h(trace_n);
                    cudaStreamSynchronize(st);
                    cudaMemcpy(h.data(), d_trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost);
                    if (FILE *f = fopen(trace_path, "w")) {
                        for (size_t w = 0; w < (size_t)nw; ++w)
                            for (size_t i = 0; i < 16; ++i) {
                                const long long *t = &h[(w * 16 + i) * 8];
                                fprintf(f, "%zu %zu", w, i + 8);
                                for (int k = 0; k < 8; ++k) fprintf(f, " %lld", t[k] ? t[k] - h[0] : -1ll);
                                fprintf(f, "\n");
                            }
                        fclose(f);
                    }
                }
                return 1;
            }
        }
    }
    // per-warp staging (two buffers per worker warp): room for ~10 bits per sample, at most the
    // worst case; a wave that outgrows it is packed straight into its record in HBM.  Short
    // waves: 12 worker warps per CTA, two CTAs per SM; longer ones: 8 worker warps (larger staging).
    const uint32_t worst = (25u * max_wave_len + 31u) / 32u + 24u;
    uint32_t stage = (10u * max_wave_len + 31u) / 32u + 24u;
    const char *e = getenv("DRICE_ENC_STAGE_WORDS");
    const bool twelve = !e && md.delta && stage <= 1120u;
    if (e) stage = (uint32_t)atol(e);
    else if (stage <= 1120u) stage = stage < 1088u ? stage : 1088u;
    else stage = 1600u;
    if (stage > worst) stage = worst;
    if (stage < 64u) stage = 64u;
    stage = (stage + 3u) & ~3u;
    auto launch = [&](auto kernel, DeviceOnce &attr_set, int nworkers) {
        const size_t smem = (size_t)stage * 2 * nworkers * sizeof(uint32_t);   // two buffers per worker warp
        const int nthreads = (nworkers + 1) * 32;
        const uint32_t ntiles = (p.nwaves + nworkers - 1) / nworkers;
        if (attr_set.first()) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        }
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, nthreads, smem);
        if (occ < 1) occ = 1;
        uint32_t grid = (uint32_t)(g_num_sms * occ);
        if (grid > ntiles) grid = ntiles;
        kernel<<<grid, nthreads, smem, st>>>(p, stage, ntiles);
    };
    static DeviceOnce attr12, attr8, attr8n;                     // per K (this function is a template)
    if (!md.delta) launch(encode_tile_kernel<K, 2, false, 8>, attr8n, 8);    // no delta (filter [1] / pre-filtered input)
    else if (twelve) launch(encode_tile_kernel<K, 2, true, 12>, attr12, 12);
    else launch(encode_tile_kernel<K, 2, true, 8>, attr8, 8);
    return 1;
}

}  // namespace

// The 16 Rice parameters are instantiated in four translation units (the same source compiled with
// DRICE_PART = 0..3, see the Makefile) so that the build parallelises; part 0 holds the dispatcher.
#ifndef DRICE_PART
#define DRICE_PART 0
#define DRICE_SINGLE_TU 1
#endif
#define DRICE_CAT2(a, b) a##b
#define DRICE_CAT(a, b) DRICE_CAT2(a, b)

int launch_encode_part0(const EncodeParams &p, const EncodeMode &m, uint32_t max_wave_len, cudaStream_t st);
int launch_encode_part1(const EncodeParams &p, const EncodeMode &m, uint32_t max_wave_len, cudaStream_t st);
int launch_encode_part2(const EncodeParams &p, const EncodeMode &m, uint32_t max_wave_len, cudaStream_t st);
int launch_encode_part3(const EncodeParams &p, const EncodeMode &m, uint32_t max_wave_len, cudaStream_t st);

#define DRICE_CASE(K) case K: return launch_k<K>(p, m, max_wave_len, st);
#ifdef DRICE_SINGLE_TU
int launch_encode_part0(const EncodeParams &p, const EncodeMode &m, uint32_t max_wave_len, cudaStream_t st)
{
    switch (p.k) {
        DRICE_CASE(0) DRICE_CASE(1) DRICE_CASE(2) DRICE_CASE(3) DRICE_CASE(4) DRICE_CASE(5)
        DRICE_CASE(6) DRICE_CASE(7) DRICE_CASE(8) DRICE_CASE(9) DRICE_CASE(10) DRICE_CASE(11)
        DRICE_CASE(12) DRICE_CASE(13) DRICE_CASE(14) DRICE_CASE(15)
    }
    return -1;
}
int launch_encode_part1(const EncodeParams &, const EncodeMode &, uint32_t, cudaStream_t) { return -1; }
int launch_encode_part2(const EncodeParams &, const EncodeMode &, uint32_t, cudaStream_t) { return -1; }
int launch_encode_part3(const EncodeParams &, const EncodeMode &, uint32_t, cudaStream_t) { return -1; }
#else
// part q holds K = 4q .. 4q+3
int DRICE_CAT(launch_encode_part, DRICE_PART)(const EncodeParams &p, const EncodeMode &m, uint32_t max_wave_len, cudaStream_t st)
{
    switch (p.k) {
        DRICE_CASE(4 * DRICE_PART) DRICE_CASE(4 * DRICE_PART + 1) DRICE_CASE(4 * DRICE_PART + 2) DRICE_CASE(4 * DRICE_PART + 3)
    }
    return -1;
}
#endif
#undef DRICE_CASE

#if DRICE_PART == 0
int launch_encode(const EncodeParams &p, const EncodeMode &m, uint32_t max_wave_len, cudaStream_t st)
{
    if (p.nwaves == 0) return 0;
    if (p.k < 0 || p.k > 15) return -1;
#ifdef DRICE_SINGLE_TU
    return launch_encode_part0(p, m, max_wave_len, st);
#else
    switch (p.k >> 2) {
        case 0: return launch_encode_part0(p, m, max_wave_len, st);
        case 1: return launch_encode_part1(p, m, max_wave_len, st);
        case 2: return launch_encode_part2(p, m, max_wave_len, st);
        default: return launch_encode_part3(p, m, max_wave_len, st);
    }
#endif
}
#endif

}  // namespace drice
