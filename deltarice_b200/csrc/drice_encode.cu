// drice_encode.cu — Delta + Rice ENCODE for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   encodeWaveform delta branch        src/deltaRice.c:49-63
//   compressWithRiceCoding             src/deltaRice.c:191-244
//   perWaveCompression                 src/deltaRice.c:365-381
//   writeWholeCompressedByteString     src/deltaRice.c:383-436 (framing + compaction)
//
// The codec is issue / integer-pipe bound on B200 long before it is HBM bound (sm_100 takes one warp
// instruction per two cycles on the ALU pipe and on the FMA pipe each), so the kernels are designed
// around instructions per sample, not bytes.
//
// encode_tile_kernel (waves of <= 8192 samples), single pass over HBM:
//   * persistent CTAs of NW worker warps + 1 control warp; a tile = NW consecutive waves; the first two
//     tiles of a CTA are fixed by its block index, later ones come from an atomic ticket taken ONE TILE
//     AHEAD, so that a worker knows its next wave while it still works on the current one; ONE WARP
//     ENCODES ONE WAVE, 16 samples (8 packed int16x2 words) per lane and round; the next round's words
//     are loaded between the round's front-end and its packing, the first round of the NEXT wave before
//     the previous wave's copy-out (the loads fly during the copy-out and the tile barrier); a worker's
//     wave state (position, size, record words) lives in shared memory between iterations, not in
//     registers - registers are what bounds the kernel's occupancy;
//   * table front-end (RiceParameter 2 ... 64; encode_round_lut): t = delta + R (R = min(4M, 32)) on packed
//     halves (PRMT / LOP3 / 2 x VIADD.16x2), the two index fields become the offset of the pair's entry
//     with LOP3 / IMAD / SHF, one LDS returns the pair's code | length << 24 from a skewed table in
//     shared memory; pairs that hold an escape (a half outside the table) are redone per sample in a
//     rare divergent path and their second code is appended after a per-position warp vote;
//   * arithmetic front-end (every other parameter, pre-filtered input, or when the table would cost
//     resident warps; encode_round): zig-zag on packed halves, two samples merged into one pair code
//     with a single IMAD;
//   * bit offsets: warp shuffle scan per round (the shuffle's own predicate guards the add), running
//     base across rounds;
//   * packing appends a code to a 64-bit window with IMAD.WIDE (acc * 2^len + value: the multiply IS
//     the shift, and it runs on the FMA pipe) and emits finished 32-bit words to the warp's staging in
//     shared memory (a word is complete when bit 5 of the lane's running bit position flips: one LOP3
//     sets the predicate of the three emit instructions); neighbouring lanes are stitched with one
//     shuffle per round (the bits a lane leaves pending are OR-ed into the first word of the next
//     lane): no shared-memory atomics;
//   * cross-wave compaction without a second pass: the worker that finishes the tile's last wave
//     publishes the tile's aggregate; the control warp resolves the tile's offset by decoupled look-back
//     while the workers encode the next tile into their second staging buffer; the record
//     [nwords][words] is copied out coalesced one iteration later, the first wave of a chunk also
//     writes the chunk header [total].  A wave larger than its staging is packed a second time straight
//     into its record.
//
// Waves longer than one tile (L > 8192): many of them - encode_multi_kernel, one CTA per wave (a sizing
// sweep, then a packing sweep that re-reads the wave and streams completed words straight to HBM); a
// few (the reference's long-wave chunkings, its default WaveformLength = -1) - the encode_long_* kernels,
// several CTAs per wave (segment sizes, one-CTA scan, packing at the segments' bit offsets).
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
    uint32_t *ptr, *first;
    uint32_t *end;      // nothing is stored at or past `end` (packing straight into HBM)

    __device__ __forceinline__ void init(uint32_t *first_word, uint32_t b0, uint32_t *end_)
    {
        n = b0 & 31u; ptr = first = first_word; end = end_; lo = 0;
    }
    __device__ __forceinline__ void store(uint32_t *q, uint32_t v) const
    {
        if (q < end) *q = v;
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
    __device__ __forceinline__ uint32_t pending() const { return n; }
    __device__ __forceinline__ bool flushed() const { return ptr != first; }      // wrote its first word itself
    __device__ __forceinline__ void store_tail(uint32_t v) const { store(ptr, v); }
};
// Packing into the warp's staging in shared memory: `pos` is the lane's running bit position (never
// wrapped: the funnel shift takes it modulo 32, and a word is complete when bit 5 of the position flips),
// the emit is four predicated instructions written in PTX so that the pointer advances in place.
template <>
struct Packer<false> {
    uint32_t lo, pos;
    uint32_t ptr, first;        // shared-memory addresses

    __device__ __forceinline__ void init(uint32_t *first_word, uint32_t b0, uint32_t *)
    {
        pos = b0 & 31u; ptr = first = (uint32_t)__cvta_generic_to_shared(first_word); lo = 0;
    }
    __device__ __forceinline__ void put(uint32_t v, uint32_t len)
    {
        const uint64_t a = mul_wide(lo, pow2(len));
        const uint32_t alo = (uint32_t)a | v, ahi = (uint32_t)(a >> 32);
        asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t, x;\n\t"
                     "add.u32 t, %1, %4;\n\t"
                     "xor.b32 x, t, %1;\n\t"
                     "and.b32 x, x, 32;\n\t"
                     "setp.ne.u32 p, x, 0;\n\t"
                     "mov.u32 %1, t;\n\t"
                     "@p shf.r.wrap.b32 x, %2, %3, t;\n\t"
                     "@p st.shared.u32 [%0], x;\n\t"
                     "@p add.u32 %0, %0, 4;\n\t}"
                     : "+r"(ptr), "+r"(pos) : "r"(alo), "r"(ahi), "r"(len) : "memory");
        lo = alo;
    }
    __device__ __forceinline__ uint32_t pending() const { return pos & 31u; }
    __device__ __forceinline__ bool flushed() const { return ptr != first; }
    __device__ __forceinline__ void store_tail(uint32_t v) const
    {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ptr), "r"(v) : "memory");
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
    if (!kFull && nvalid == 0) pw = 0;                   // lanes past the wave's end code nothing: no escape at the boundary
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
    uint32_t *const first = dst + (b0 >> 5);
    pk.init(first, b0, dst + cap);
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
    asm("shl.b32 %0, %1, %2;" : "=r"(frag) : "r"(pk.lo), "r"(32u - pk.pending()));   // nothing pending -> 0
    if (!packs) frag = 0;
    const bool flushed = pk.flushed();                   // wrote its first word itself
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
        if (pk.pending()) pk.store_tail(frag);
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

// ======================================================================================
// table front-end: one shared-memory lookup per PAIR of samples
// ======================================================================================
// For 1 <= K <= 6 the Rice split of a pair of deltas (src/deltaRice.c:207-222) is a table: when both
// deltas lie in [-R, R) with R = min(4M, 32) (quotients < 8, no escape), the pair's code and length are a
// function of 2(K+3) bits.  The lane forms t = delta + R on packed halves (no zig-zag: the table
// absorbs it), one LOP3 / IMAD / SHF turn the two (K+3)-bit fields into the entry's offset and one
// LDS fetches code | length << 24.  Halves outside [0, 2R) flag the pair: it is redone per sample
// (escape rule, :223-228) in a rare divergent path, as in the arithmetic front-end.
// Entry (lo, hi) sits at word lo + hi * (2^(K+3) + kSkew): deltas cluster around 0, so without the
// skew the bank (= lo mod 32) would be the same handful for every lane.
template <int K>
struct LutConst {
    static constexpr bool     kOk   = (K >= 1 && K <= 6);
    static constexpr uint32_t IDXB  = (K + 3 < 6) ? K + 3 : 6;                 // bits of one biased delta (table of <= 17.6 KB)
    static constexpr uint32_t R     = 1u << (IDXB - 1);                        // 4M for K <= 3, 32 (quotients < 64 / 2M) above
    static constexpr uint32_t kSkew = 5;
    static constexpr uint32_t ROW   = (1u << IDXB) + kSkew;
    static constexpr uint32_t LOW   = (2u * R - 1u) * 0x10001u;
    static constexpr uint32_t HIGH  = ~LOW;
    static constexpr uint32_t MULT  = 0x10000u + ROW;                          // (t & LOW) * MULT: the entry's index at bit 16
    static constexpr uint32_t ENTRIES = (2u * R - 1u) * ROW + 2u * R;
};
template <int K>
__device__ __forceinline__ void build_pair_table(uint32_t *tab, int tid, int nthreads)
{
    using C = LutConst<K>;
    for (uint32_t i = tid; i < 4u * C::R * C::R; i += nthreads) {
        const uint32_t tl = i & (2u * C::R - 1u), th = i >> C::IDXB;
        const int dl = (int)tl - (int)C::R, dh = (int)th - (int)C::R;
        const uint32_t ul = dl >= 0 ? 2u * dl : (uint32_t)(-2 * dl - 1), uh = dh >= 0 ? 2u * dh : (uint32_t)(-2 * dh - 1);
        uint32_t vl, ll, vh, lh;
        rice_code<K>(ul, vl, ll);
        rice_code<K>(uh, vh, lh);
        tab[tl + th * C::ROW] = ((vl << lh) | vh) | ((ll + lh) << 24);   // the first sample (low half) leads
    }
}
__device__ __forceinline__ uint32_t lds32a(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// one step of a warp inclusive scan: the shuffle's own predicate (source lane in range) guards the add
__device__ __forceinline__ uint32_t scan_up(uint32_t v, int d)
{
    uint32_t r;
    asm volatile("{ .reg .u32 r0; .reg .pred p; shfl.sync.up.b32 r0|p, %1, %2, 0, 0xffffffff; @p add.u32 r0, r0, %1; mov.u32 %0, r0; }"
                 : "=r"(r) : "r"(v), "r"(d));
    return r;
}

// NW8 packed words (2 * NW8 consecutive samples) of one lane; same alignment rules as load_slot
template <int NW8>
__device__ __forceinline__ void load_words(uint32_t (&w)[NW8], const int16_t *q, uint32_t mis)
{
    if (mis == 0) {
#pragma unroll
        for (int v = 0; v < NW8 / 4; ++v) {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(q) + v);
            w[4 * v] = a.x; w[4 * v + 1] = a.y; w[4 * v + 2] = a.z; w[4 * v + 3] = a.w;
        }
    } else if (mis == 4) {
#pragma unroll
        for (int v = 0; v < NW8 / 2; ++v) {
            const uint2 a = __ldg(reinterpret_cast<const uint2 *>(q) + v);
            w[2 * v] = a.x; w[2 * v + 1] = a.y;
        }
    } else if ((mis & 1) == 0) {
#pragma unroll
        for (int v = 0; v < NW8; ++v) w[v] = __ldg(reinterpret_cast<const uint32_t *>(q) + v);
    } else {
        const uint32_t *qa = reinterpret_cast<const uint32_t *>(q - 1);
        uint32_t t[NW8 + 1];
        t[0] = (uint32_t)(uint16_t)q[0] << 16;
#pragma unroll
        for (int v = 1; v < NW8; ++v) t[v] = __ldg(qa + v);
        t[NW8] = (uint32_t)(uint16_t)q[2 * NW8 - 1];
#pragma unroll
        for (int v = 0; v < NW8; ++v) w[v] = prmt(t[v], t[v + 1], 0x5432);
    }
}
template <int NW8>
__device__ __forceinline__ void load_words_tail(uint32_t (&w)[NW8], const int16_t *q, uint32_t mis, uint32_t nvalid,
                                                const int16_t *hi)
{
#pragma unroll
    for (int m = 0; m < NW8; ++m) w[m] = 0;
    if (nvalid == 0) return;
    if (q + 2 * NW8 <= hi) { load_words<NW8>(w, q, mis); return; }
#pragma unroll
    for (int i = 0; i < 2 * NW8; ++i)
        if (q + i < hi) w[i >> 1] |= (uint32_t)(uint16_t)q[i] << (16 * (i & 1));
}

// t = delta + R on packed halves: w - sh = w + ~sh + 1 with sh = (low half of w << 16) | high half of prev
template <int K, bool kDelta>
__device__ __forceinline__ uint32_t biased_delta(uint32_t w, uint32_t prev)
{
    using C = LutConst<K>;
    if (!kDelta) return __vadd2(w, C::R * 0x10001u);
    return __vadd2(__vadd2(w, prmt(~prev, ~w, 0x5432)), (C::R + 1u) * 0x10001u);
}

// One round of the table front-end: 2 * NW8 samples per lane.  `prefetch` is called between the
// front-end (which consumes w) and the scan / pack phase: the caller loads the NEXT round's words
// into w there, so the loads fly while this round is packed and no register copies are needed.
template <int K, int NW8, bool kDirect, bool kFull, bool kDelta, class Prefetch>
__device__ __forceinline__ void encode_round_lut(uint32_t (&w)[NW8], uint32_t nvalid, bool last_lane, int lane,
                                                 uint32_t *dst, uint32_t cap, SweepState &st, uint32_t tab,
                                                 Prefetch &&prefetch)
{
    using C = LutConst<K>;
    constexpr uint32_t SS = 2u * NW8;
    if (!kFull && nvalid > 0 && nvalid < SS) {           // short last slot: see encode_round
        const uint32_t sel_hi = kDelta ? 0x1010u : 0x4410u, sel_all = kDelta ? 0x3232u : 0x4444u;
#pragma unroll
        for (int v = 0; v < NW8; ++v) {
            if (2u * v + 1 == nvalid) w[v] = prmt(w[v], 0, sel_hi);
            if (v > 0 && 2u * v >= nvalid) w[v] = prmt(w[v - 1], 0, sel_all);
        }
    }
    uint32_t pw = __shfl_up_sync(0xffffffffu, w[NW8 - 1], 1);
    if (lane == 0) pw = st.prev_last;
    if (!kFull && nvalid == 0) pw = 0;                   // lanes past the wave's end code nothing: their (zero) words must not flag an escape at the boundary
    st.prev_last = __shfl_sync(0xffffffffu, w[NW8 - 1], 31);

    // ---- one lookup per pair ------------------------------------------------------------------
    uint32_t ev[NW8], t[NW8];
    uint32_t uor = 0;
    {
        uint32_t np = ~pw;
#pragma unroll
        for (int m = 0; m < NW8; ++m) {
            if (kDelta) {
                const uint32_t nw = ~w[m];
                t[m] = __vadd2(__vadd2(w[m], prmt(np, nw, 0x5432)), (C::R + 1u) * 0x10001u);
                np = nw;
            } else {
                t[m] = __vadd2(w[m], C::R * 0x10001u);
            }
            uor |= t[m];
            ev[m] = lds32a(tab + (((t[m] & C::LOW) * C::MULT) >> 14));   // (the product's bits 14, 15 are zero)
        }
    }
    // lengths sit in byte 3 and the codes stay below 2^20: eight entries add up without a carry into it
    uint32_t T = 0;
#pragma unroll
    for (int h = 0; h < NW8 / 8; ++h) {
        uint32_t sacc = 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) sacc += ev[8 * h + m];
        T += sacc >> 24;
    }
    // A pair with a half outside the table is coded per sample (escape rule): its first sample's code
    // replaces the entry, the second one's goes to sec[m] (0 = nothing: a zero-length append is a no-op),
    // both as value | length << 24 (an escape's value has 17 bits)
    const bool flagged = (uor & C::HIGH) != 0u;
    const bool any_flag = __any_sync(0xffffffffu, flagged);
    if (any_flag) {
#pragma unroll
        for (int m = 0; m < NW8; ++m) {
            const uint32_t tt = t[m];
            t[m] = 0;                                    // from here on t[m] is sec[m]
            if (tt & C::HIGH) {
                // zig-zag of the two deltas (t - R), then the escape-aware code of each
                const uint32_t d2 = __vsub2(tt, C::R * 0x10001u);
                const uint32_t U = __vadd2(d2, d2) ^ prmt(d2, 0, 0xbb99);
                uint32_t v0, l0, v1, l1;
                rice_code<K>(U & 0xFFFFu, v0, l0);
                rice_code<K>(U >> 16, v1, l1);
                T += l0 + l1 - (ev[m] >> 24);
                ev[m] = v0 | (l0 << 24);
                t[m] = v1 | (l1 << 24);
            }
        }
    }
    if (!kFull && nvalid < SS) T = nvalid ? T - (SS - nvalid) * (K + 1) : 0u;
    prefetch();

    // ---- warp exclusive scan of T ----------------------------------------------------------
    uint32_t inc = T;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) inc = scan_up(inc, d);
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    const uint32_t b0 = st.base + inc - T;
    st.base += total;
    if (!kDirect && st.base > cap * 32u - 544u) st.ovf = true;     // (the round after this one may not fit)
    if (st.ovf) return;

    // ---- pack ----------------------------------------------------------------------------------
    Packer<kDirect> pk;
    uint32_t *const first = dst + (b0 >> 5);
    pk.init(first, b0, dst + cap);
    const bool packs = kFull || nvalid > 0;
    if (packs) {
        if (!any_flag) {
#pragma unroll
            for (int m = 0; m < NW8; ++m) pk.put(ev[m] & 0xFFFFFFu, ev[m] >> 24);
        } else {
            const uint32_t amask = kFull ? 0xffffffffu : __activemask();
#pragma unroll
            for (int m = 0; m < NW8; ++m) {
                pk.put(ev[m] & 0xFFFFFFu, ev[m] >> 24);
                if (__any_sync(amask, t[m] != 0u)) pk.put(t[m] & 0xFFFFFFu, t[m] >> 24);
            }
        }
    }
    // ---- stitch the lanes (see encode_round) ---------------------------------------------------
    uint32_t frag;
    asm("shl.b32 %0, %1, %2;" : "=r"(frag) : "r"(pk.lo), "r"(32u - pk.pending()));
    if (!packs) frag = 0;
    uint32_t from_prev = __shfl_up_sync(0xffffffffu, frag, 1);
    if (lane == 0) from_prev = st.carry_round;
    if (kFull && !kDirect) {
        // a full lane holds >= 16 (K + 1) >= 32 bits: its first word is in the staging already
        *first |= from_prev;
    } else if (packs) {
        const bool flushed = pk.flushed();
        if (flushed) { if (from_prev) { if (!kDirect || first < dst + cap) *first |= from_prev; } }
        else frag |= from_prev;
    }
    st.carry_round = __shfl_sync(0xffffffffu, frag, 31);
    if (!kFull && last_lane && packs) {
        if (pk.pending()) pk.store_tail(frag);
        if (st.base & 31u) dst[st.base >> 5] &= 0xFFFFFFFFu << (32u - (st.base & 31u));
    }
}

// Sweep over one wave with the table front-end: rounds of 32 * 2 * NW8 samples while they are full,
// then at most 2 * NW8 / 16 generic rounds of 512 (the wave's last lane must close the final word in one).
template <int K, int NW8, bool kDirect, bool kDelta>
__device__ __forceinline__ uint32_t encode_wave_lut(const int16_t *wave, uint32_t n, const int16_t *raw_hi, int lane,
                                                    uint32_t *dst, uint32_t cap, bool *overflow, uint32_t tab,
                                                    uint32_t (&w)[NW8], bool preloaded)
{   // preloaded (warp uniform): w holds the first full round's words already (requested before the previous
    // wave's copy-out, see the tile kernel)
    static_assert(NW8 == 8, "one generic round of 512 samples closes the wave");
    constexpr uint32_t kBig = 64u * NW8;                 // samples per full round
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(wave) & 15u) >> 1);
    const uint32_t nbig = n ? (n - 1u) / kBig : 0u;      // the rest (1 .. kBig samples) goes to the generic round
    SweepState st;
    st.base = 0;
    st.carry_round = 0;
    st.prev_last = 0;
    st.ovf = false;
    // the generic round's slot of this lane
    const uint32_t s0 = nbig * kBig + lane * 16u;
    const int32_t rem = (int32_t)n - (int32_t)s0;
    const uint32_t nv = rem >= 16 ? 16u : (rem > 0 ? (uint32_t)rem : 0u);
    if (nbig) {
        const int16_t *q = wave + lane * (2 * NW8);
        if (!preloaded) load_words<NW8>(w, q, mis);
        for (uint32_t r = 0; r < nbig; ++r) {
            q += kBig;
            // the next round's words (the last full round: the generic round's) fly while this one is packed
            encode_round_lut<K, NW8, kDirect, true, kDelta>(w, 2u * NW8, false, lane, dst, cap, st, tab, [&] {
                if (r + 1 < nbig) load_words<NW8>(w, q, mis);
                else load_words_tail<NW8>(w, wave + s0, mis, nv, raw_hi);
            });
        }
    } else {
        load_words_tail<NW8>(w, wave + s0, mis, nv, raw_hi);
    }
    encode_round_lut<K, NW8, kDirect, false, kDelta>(w, nv, nv > 0 && s0 + 16u >= n, lane, dst, cap, st, tab, [] {});
    __syncwarp();
    *overflow = st.ovf;
    return st.base;
}
// the rare wave that outgrew its staging: out of line, so that it stays out of the hot code
template <int K, int NW8, bool kDelta>
__device__ __noinline__ void encode_wave_lut_in_place(const int16_t *wave, uint32_t n, const int16_t *raw_hi, int lane,
                                                      uint32_t *dst, uint32_t cap, uint32_t tab)
{
    bool dummy;
    uint32_t w[NW8];
    encode_wave_lut<K, NW8, true, kDelta>(wave, n, raw_hi, lane, dst, cap, &dummy, tab, w, false);
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
// Shared control state lives in a ring of kRing slots (iteration % kRing): a slot is reused only after
// the workers have copied out the tile of that slot, which needs the control warp to be done with it.
constexpr int kRing = 4;                                  // per-tile state: previous / current / next + one a fast worker has moved on to
constexpr int kTileRing = 8;                              // tile indices, published three iterations ahead

// NW worker warps (+ 1 control warp) per CTA: 24 in one CTA per SM (best since the workers run without a
// barrier between them), 12 in two CTAs or 8 when the staging of longer records needs the room, 8 in two
// CTAs for small batches (launch_k)
// LUT: 0 = arithmetic front-end (encode_round), 1 = table front-end (1 <= K <= 6), 16 samples per lane and round
// (32 per lane were measured too: twice as slow - registers spill and the unrolled round outgrows the
// instruction cache)
template <int LUT> struct LutVariant { static constexpr int NW8 = 8; };
template <int K, bool kDirect, bool kDelta, int LUT>
__device__ __forceinline__ uint32_t encode_wave_any(const int16_t *wave, uint32_t n, const int16_t *raw_hi, int lane,
                                                    uint32_t *dst, uint32_t cap, bool *overflow, uint32_t tab,
                                                    uint32_t (&w)[8], bool preloaded)
{
    if constexpr (LUT == 0) {
        return encode_wave<K, kDirect, kDelta>(wave, n, raw_hi, lane, dst, cap, overflow);
    } else if constexpr (kDirect) {
        encode_wave_lut_in_place<K, LutVariant<LUT>::NW8, kDelta>(wave, n, raw_hi, lane, dst, cap, tab);
        return 0;
    } else {
        return encode_wave_lut<K, LutVariant<LUT>::NW8, false, kDelta>(wave, n, raw_hi, lane, dst, cap, overflow, tab, w, preloaded);
    }
}
// A worker's wave, kept in shared memory between the iterations of the tile kernel (set up one iteration
// ahead, encoded, copied out one iteration later): registers are what limits the kernel's occupancy.
struct WaveSlot {
    uint64_t begin;         // first sample of the wave in raw
    uint32_t n;             // samples
    uint32_t chunk;
    uint32_t first;         // 1 if first wave of its chunk
    uint32_t chunk_total;
    uint32_t nwords;        // record words (after encoding)
    uint32_t flags;         // bit 0: there is a wave, bit 1: it outgrew its staging
};
// Geometry of the wave a worker takes from `tile`; for the table front-end the first full round's words are
// requested here (waves of more than one round).
template <int NW, int LUT>
__device__ __forceinline__ void setup_wave(const EncodeParams &p, uint32_t tile, uint32_t ntiles, int warp, int lane,
                                           WaveSlot *slot, uint32_t (&w)[8], bool &preloaded)
{
    preloaded = false;
    const uint32_t g = tile * NW + warp;
    if (tile >= ntiles || g >= p.nwaves) {
        if (lane == 0) slot->flags = 0;
        return;
    }
    const WaveGeom wg = locate_wave(p, g);
    if (lane == 0) {
        slot->begin = wg.begin; slot->n = wg.n; slot->chunk = wg.chunk; slot->first = wg.first;
        slot->chunk_total = wg.chunk_total; slot->nwords = 0; slot->flags = 1;
    }
    if (LUT != 0 && wg.chunk_total && wg.n > 512u) {
        const int16_t *wave = p.raw + wg.begin;
        load_words<8>(w, wave + lane * 16, (uint32_t)((reinterpret_cast<uintptr_t>(wave) & 15u) >> 1));
        preloaded = true;
    }
}
template <int K, int LUT>
constexpr size_t lut_table_bytes() { return LUT == 0 ? 0 : (size_t)((LutConst<K>::ENTRIES + 3u) & ~3u) * 4; }

template <int K, int MINB, bool kDelta, int NW, int LUT = 0>
__global__ void __launch_bounds__((NW + 1) * 32, MINB)
encode_tile_kernel(const EncodeParams p, const uint32_t stage_words, const uint32_t ntiles, uint32_t *const max_words)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ volatile uint32_t s_tile[kTileRing];      // tile index of the iteration
    __shared__ volatile uint32_t s_tseq[kTileRing];      // = iteration + 1 once s_tile holds that iteration's tile
    __shared__ uint32_t s_mine[kRing][NW];        // words each wave contributes
    __shared__ uint32_t s_cnt[kRing];                    // workers that have reported
    __shared__ uint32_t s_total[kRing];
    __shared__ uint64_t s_off[kRing];                    // tile's exclusive word offset
    __shared__ volatile uint32_t s_flag[kRing];          // = it + 1 once s_off is valid
    __shared__ WaveSlot s_wave[kRing][NW];               // the workers' waves: next / current / previous
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool control = warp == NW;

    if (threadIdx.x < kRing) { s_flag[threadIdx.x] = 0; s_cnt[threadIdx.x] = 0; }
    // the first two tiles of a CTA are fixed (all CTAs are resident and run iteration 0 together, so tiles
    // still start in index order); later ones come from the ticket counter, one iteration ahead
    if (threadIdx.x < kTileRing) {
        s_tile[threadIdx.x] = threadIdx.x * gridDim.x + blockIdx.x;
        s_tseq[threadIdx.x] = threadIdx.x < 3 ? threadIdx.x + 1 : 0;
    }
    uint32_t tab = 0;                                    // shared address of the pair table (behind the staging)
    if constexpr (LUT != 0) {
        uint32_t *const tabp = smem + (size_t)(2 * NW) * stage_words;
        build_pair_table<K>(tabp, threadIdx.x, (NW + 1) * 32);
        tab = (uint32_t)__cvta_generic_to_shared(tabp);
    }
    __syncthreads();

    if (control) {
        for (uint32_t it = 0;; ++it) {
            const int slot = it % kRing;
            // sleeps on a named barrier (one per ring slot) until the tile's last worker arrives
            asm volatile("bar.sync %0, 64;" ::"r"(2 + slot) : "memory");
            const uint32_t tile = s_tile[it % kTileRing];
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
    // Tile indices are published ahead of their iteration (s_tile / s_tseq: the first three of a CTA are fixed,
    // later ones are tickets warp 0 takes during iteration it for iteration it + 3), so a worker sets up its
    // next wave - geometry and, for the table front-end, the first round's words - before it copies out the
    // previous one: the loads fly during the copy-out.
    uint32_t *const stage0 = smem + (size_t)(2 * warp) * stage_words;   // two staging buffers per warp
    const int16_t *const raw_hi = p.raw + p.raw_samples;
    uint32_t largest = 0;                                // largest record of this warp (sizes the next batch's staging)
    uint32_t w0[8];                                      // first round of the wave set up ahead
    bool pre = false;
    setup_wave<NW, LUT>(p, s_tile[0], ntiles, warp, lane, &s_wave[0][warp], w0, pre);
    __syncwarp();
    for (uint32_t it = 0;; ++it) {
        const int par = it & 1, slot = it % kRing;
        const uint32_t tile = s_tile[it % kTileRing];
        const bool live = tile < ntiles;
        uint32_t next_ticket = 0;
        bool pre_next = false;
        if (live) {
            uint32_t mine = 0;
            WaveSlot *const ws = &s_wave[slot][warp];
            if (ws->flags) {
                if (ws->chunk_total) {
                    bool ovf = false;
                    const uint32_t nwords = (encode_wave_any<K, false, kDelta, LUT>(p.raw + ws->begin, ws->n, raw_hi, lane, stage0 + par * stage_words,
                                                                            stage_words, &ovf, tab, w0, pre) + 31u) >> 5;
                    mine = nwords + 1u;
                    largest = nwords > largest ? nwords : largest;
                    if (lane == 0) { ws->nwords = nwords; ws->flags = ovf ? 3u : 1u; }
                }
                mine += ws->first;                           // empty chunk: header only
            }
            // the tile after the next: taken as late as possible so that tiles start in ticket order; the
            // atomic's latency hides behind the copy-out below
            if (threadIdx.x == 0) next_ticket = 3u * gridDim.x + atomicAdd(p.ticket, 1u);
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
            // (the next tile's index was published two iterations ago by warp 0, which may lag this warp by one)
            while (s_tseq[(it + 1) % kTileRing] != it + 2) __nanosleep(20);
            setup_wave<NW, LUT>(p, s_tile[(it + 1) % kTileRing], ntiles, warp, lane, &s_wave[(it + 1) % kRing][warp], w0, pre_next);
        } else if (warp == 0) {
            __threadfence_block();
            asm volatile("bar.arrive %0, 64;" ::"r"(2 + slot) : "memory");   // lets the control warp see the end
        }
        // ---- copy out the wave of the previous iteration ------------------------------------
        const int ps = (it + kRing - 1) % kRing;
        const WaveSlot *const wp = &s_wave[ps][warp];
        if (it > 0 && wp->flags) {
            while (s_flag[ps] != it) __nanosleep(100);        // tile offset: normally there long ago
            __threadfence_block();
            const uint32_t v = lane < NW ? s_mine[ps][lane] : 0u;
            const uint32_t loff = __reduce_add_sync(0xffffffffu, lane < warp ? v : 0u);
            const uint64_t off = s_off[ps] + loff;
            const uint32_t nwords_prev = wp->nwords, first_prev = wp->first, total_prev = wp->chunk_total;
            const uint32_t rec_words = total_prev ? nwords_prev + 1u : 0u;
            const bool fits = off + rec_words + first_prev <= p.out_cap_words;
            if (lane == 0 && first_prev) p.chunk_byte_off[wp->chunk] = off * 4;
            if (fits) {
                uint32_t *rec = p.out + off + first_prev;
                if (lane == 0) {
                    if (first_prev) p.out[off] = total_prev;
                    if (rec_words) rec[0] = nwords_prev;
                }
                if (rec_words) {
                    if (!(wp->flags & 2u)) {
                        const uint32_t *src = stage0 + (par ^ 1) * stage_words;
                        uint32_t i = lane;
                        for (; i + 96u < nwords_prev; i += 128u) {          // four independent words per lane and trip
                            const uint32_t a = src[i], b = src[i + 32], c = src[i + 64], d = src[i + 96];
                            rec[1 + i] = a; rec[33 + i] = b; rec[65 + i] = c; rec[97 + i] = d;
                        }
                        for (; i < nwords_prev; i += 32u) rec[1 + i] = src[i];
                    } else {                                 // larger than the staging: pack in place
                        bool dummy;
                        uint32_t wd[8];
                        encode_wave_any<K, true, kDelta, LUT>(p.raw + wp->begin, wp->n, raw_hi, lane, rec + 1, nwords_prev, &dummy, tab, wd, false);
                    }
                }
            }
            __syncwarp();
        }
        if (!live) break;
        if (threadIdx.x == 0) {
            s_tile[(it + 3) % kTileRing] = next_ticket;
            __threadfence_block();
            s_tseq[(it + 3) % kTileRing] = it + 4;
        }
        pre = pre_next;
        __syncwarp();                                        // lane 0's wave slot, for the lanes of the next iteration
        // No barrier between the workers: a worker moves on to its next wave at once.  What orders them is the
        // copy-out above - it needs the tile's offset, i.e. every worker's wave of the PREVIOUS iteration - so a
        // worker is never more than one iteration ahead of the slowest one (kRing = 4 slots of tile state).
    }
    if (lane == 0 && max_words && largest) atomicMax(max_words, largest);
}

// zig-zag of the 16-bit wrapped difference cur-prev (src/deltaRice.c:57-62, :207-211)
__device__ __forceinline__ uint32_t zigzag_delta(int cur, int prev)
{
    const int t = cur - prev;
    const uint32_t s = (uint32_t)((int)((uint32_t)t << 16) >> 31);
    return (((uint32_t)t << 1) ^ s) & 0xFFFFu;
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

// ======================================================================================
// long waves in small batches: several CTAs per wave
// ======================================================================================
// One CTA per wave (encode_multi_kernel) leaves most SMs idle when a batch has a handful of very long waves
// - the reference's long-wave configurations (docs/Performance.md:27-47) and its DEFAULT WaveformLength = -1
// (the whole chunk is one wave).  Here a wave is cut into segments of kLongSegSlots slots (16 samples each,
// aligned to memory as in slot_codes); work item = (wave, segment):
//   1. encode_long_size_kernel  - bits of every segment (a sizing sweep);
//   2. encode_long_scan_kernel  - one CTA: per wave (a warp each) the segments' exclusive bit offsets and the
//      record's word count, then across the waves the record offsets; writes the chunk / record headers and
//      the chunk byte offsets, and zeroes the words two segments share;
//   3. encode_long_pack_kernel  - packs every segment at its bit offset: whole words with plain stores, the
//      first and last (shared) word with atomicOr.
constexpr uint32_t kLongSegTiles = 2;                                  // tiles of kEncMaxThreads slots per segment
constexpr uint32_t kLongSegSlots = kLongSegTiles * kEncMaxThreads;     // 16384 samples

struct LongParams {
    EncodeParams p;
    uint64_t *seg_bits;      // [nwaves * maxseg]: bits of a segment, then (after the scan) its exclusive bit offset
    uint64_t *wave_off;      // [nwaves]: word offset of the wave's first output word (chunk header or record header)
    uint32_t  maxseg;
    int       dmask;
};

__device__ __forceinline__ uint32_t long_wave_slots(const int16_t *wave, uint32_t n)
{
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(wave) & 31u) >> 1);
    return (uint32_t)(((uint64_t)mis + n + S - 1) / S);
}

template <int K>
__global__ void __launch_bounds__(kEncMaxThreads)
encode_long_size_kernel(const LongParams lp)
{
    __shared__ uint64_t s64[kEncMaxThreads / 32];
    const EncodeParams &p = lp.p;
    const uint32_t g = blockIdx.y, seg = blockIdx.x;
    const WaveGeom wg = locate_wave(p, g);
    const int16_t *wave = p.raw + wg.begin;
    const uint32_t nslots = long_wave_slots(wave, wg.n);
    const uint32_t s_lo = seg * kLongSegSlots;
    if (s_lo >= nslots) return;
    const uint32_t s_hi = min(nslots, s_lo + kLongSegSlots);
    const int mis = (int)((reinterpret_cast<uintptr_t>(wave) & 31u) >> 1);
    uint32_t cv[S];
    uint64_t bits = 0;
    for (uint32_t sl = s_lo + threadIdx.x; sl < s_hi; sl += kEncMaxThreads)
        bits += slot_codes<K>(wave, (int64_t)sl * S - mis, wg.n, cv, lp.dmask);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, d);
    if ((threadIdx.x & 31) == 0) s64[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t tot = 0;
        for (int w = 0; w < kEncMaxThreads / 32; ++w) tot += s64[w];
        lp.seg_bits[(size_t)g * lp.maxseg + seg] = tot;
    }
}

__global__ void __launch_bounds__(1024)
encode_long_scan_kernel(const LongParams lp)
{
    const EncodeParams &p = lp.p;
    __shared__ uint64_t s_words[1024];       // output words of each wave (record + chunk header), then their exclusive scan
    __shared__ uint64_t s_warp[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ---- per wave (a warp each): exclusive bit offsets of its segments, record size --------------------
    for (uint32_t g = warp; g < p.nwaves; g += 32) {
        const WaveGeom wg = locate_wave(p, g);
        const uint32_t nslots = long_wave_slots(p.raw + wg.begin, wg.n);
        const uint32_t nseg = (nslots + kLongSegSlots - 1) / kLongSegSlots;
        uint64_t *sb = lp.seg_bits + (size_t)g * lp.maxseg;
        uint64_t run = 0;
        for (uint32_t s0 = 0; s0 < nseg; s0 += 32) {
            const uint32_t sg = s0 + lane;
            const uint64_t v = sg < nseg ? sb[sg] : 0ull;
            uint64_t inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (sg < nseg) sb[sg] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            const uint64_t nwords = (run + 31u) >> 5;
            // the wave's words: [chunk total] (first wave of a chunk) + [nwords] + words (nothing for an empty chunk)
            s_words[g] = (wg.chunk_total ? nwords + 1u : 0u) + wg.first;
            // total bits kept for the pack kernel's end word: slot nseg of the row (maxseg has room for it)
            sb[nseg] = run;
        }
    }
    __syncthreads();
    // ---- across the waves: exclusive scan of their words (nwaves <= 1024) -----------------------------
    {
        const uint32_t g = threadIdx.x;
        const uint64_t v = g < p.nwaves ? s_words[g] : 0ull;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint64_t ws = s_warp[lane];
            uint64_t wi = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += t;
            }
            s_warp[lane] = wi - ws;
            if (lane == 31) s_warp[32] = wi;
        }
        __syncthreads();
        const uint64_t excl = s_warp[warp] + inc - v;
        const uint64_t total = s_warp[32];
        const bool fits = total <= p.out_cap_words;
        if (threadIdx.x == 0) {
            if (!fits) atomicOr(p.status, kErrCapacity);
            p.chunk_byte_off[p.nchunks] = total * 4;
        }
        __syncthreads();
        if (g < p.nwaves) {
            s_words[g] = excl;
            lp.wave_off[g] = fits ? excl : ~0ull;                // ~0: nothing is written
        }
        __syncthreads();
        if (!fits) {
            // offsets are still reported (as the other kernels do); no data is written
            if (g < p.nwaves) {
                const WaveGeom wg = locate_wave(p, g);
                if (wg.first) p.chunk_byte_off[wg.chunk] = excl * 4;
            }
            return;
        }
    }
    // ---- headers, and zeroes under every word two segments share (a warp per wave) --------------------
    for (uint32_t g = warp; g < p.nwaves; g += 32) {
        const WaveGeom wg = locate_wave(p, g);
        const uint64_t off = s_words[g];
        const uint32_t nslots = long_wave_slots(p.raw + wg.begin, wg.n);
        const uint32_t nseg = (nslots + kLongSegSlots - 1) / kLongSegSlots;
        const uint64_t *sb = lp.seg_bits + (size_t)g * lp.maxseg;
        uint32_t *rec = p.out + off + wg.first;
        if (lane == 0) {
            if (wg.first) { p.out[off] = wg.chunk_total; p.chunk_byte_off[wg.chunk] = off * 4; }
            if (wg.chunk_total) rec[0] = (uint32_t)((sb[nseg] + 31u) >> 5);
        }
        if (wg.chunk_total) {
            for (uint32_t sg = lane; sg <= nseg; sg += 32) {     // (entry nseg = the wave's end)
                const uint64_t b = sb[sg];
                if (b & 31u) rec[1 + (b >> 5)] = 0u;
            }
        }
    }
}

// packs the slots [s_lo, s_hi) of one wave; P0 = bit offset of the segment's first bit in the record
template <int K>
__device__ __forceinline__ void pack_segment_streaming(const int16_t *wave, uint32_t n, uint32_t *rec, uint32_t *smem, int dmask,
                                                       uint32_t s_lo, uint32_t s_hi, uint64_t P0)
{
    constexpr int NT = kEncMaxThreads;
    const int tid = threadIdx.x;
    uint32_t *sbits = smem;                  // NT*13 + 1 (worst case 12.5 words per slot)
    uint32_t *shead = sbits + NT * 13 + 1;   // NT
    uint32_t *sboff = shead + NT;            // NT
    uint32_t *swarp = sboff + NT;            // 33
    __shared__ uint32_t s_carry;
    const int mis = (int)((reinterpret_cast<uintptr_t>(wave) & 31u) >> 1);
    uint32_t cv[S];
    if (tid == 0) s_carry = 0;
    __syncthreads();
    uint64_t P = P0;                           // bits of the record before the current tile
    const uint64_t shared_first = (P0 & 31u) ? (P0 >> 5) : ~0ull;   // word shared with the previous segment
    for (uint32_t t0 = s_lo; t0 < s_hi; t0 += NT) {
        const uint32_t sl = t0 + tid;
        uint32_t T = 0;
        if (sl < s_hi) T = slot_codes<K>(wave, (int64_t)sl * S - mis, n, cv, dmask);
        else {
#pragma unroll
            for (int j = 0; j < S; ++j) cv[j] = 0;
        }
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
        if (tid == 0 && pre) {                 // word 0 was started by the previous tile (or segment)
            uint32_t c = s_carry;
            for (uint32_t j = 0; j < (uint32_t)NT && (sboff[j] >> 5) == 0; ++j) c |= shead[j];
            sbits[0] = c;
        }
        __syncthreads();
        const uint32_t bend = pre + total;
        const uint32_t full = bend >> 5;
        const uint64_t w0 = P >> 5;
        uint32_t *dstw = rec + 1 + w0;
        for (uint32_t w = tid; w < full; w += NT) {
            if (w0 + w == shared_first) atomicOr(dstw + w, sbits[w]);
            else dstw[w] = sbits[w];
        }
        if (tid == 0) s_carry = (bend & 31u) ? sbits[full] : 0u;
        P += total;
        __syncthreads();
    }
    if (tid == 0 && (P & 31u)) atomicOr(rec + 1 + (P >> 5), s_carry);   // shared with the next segment (zeroed by the scan)
}

template <int K>
__global__ void __launch_bounds__(kEncMaxThreads)
encode_long_pack_kernel(const LongParams lp)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const EncodeParams &p = lp.p;
    const uint32_t g = blockIdx.y, seg = blockIdx.x;
    const WaveGeom wg = locate_wave(p, g);
    if (wg.chunk_total == 0) return;
    const int16_t *wave = p.raw + wg.begin;
    const uint32_t nslots = long_wave_slots(wave, wg.n);
    const uint32_t s_lo = seg * kLongSegSlots;
    if (s_lo >= nslots) return;
    const uint64_t off = lp.wave_off[g];
    if (off == ~0ull) return;                                    // does not fit: flagged by the scan
    uint32_t *rec = p.out + off + wg.first;
    pack_segment_streaming<K>(wave, wg.n, rec, smem, lp.dmask, s_lo, min(nslots, s_lo + kLongSegSlots),
                              lp.seg_bits[(size_t)g * lp.maxseg + seg]);
}

// Launch geometry of the tile kernel.  Shared memory per CTA = two staging buffers of `stage` words per
// worker warp (+ the pair table); a wave that outgrows its staging is packed a second time straight into
// its record, so the staging should hold the batch's largest record: it is sized from the largest record
// of the context's previous batch of the same shape (`words_hint`, + 1/16), else for 10 bits per sample.
// More resident worker warps beat larger staging (measured with a barrier per tile, L = 7000: 2 x 12 warps
// 0.44 ms, 1 x 24 0.45, 1 x 20 0.48, 1 x 12 0.65, 1 x 8 0.90; without the barrier 1 x 24 is the best: C2 0.530
// against 0.559 for 2 x 12), so the geometries 1 x 24, 2 x 12, 2 x 8, 1 x 8 are tried in that order and the
// first one whose staging holds `want` words wins; without a hint the first geometry is taken with all the
// staging it has room for.  Small batches take 8-wave tiles (more SMs, less contention per wave).
struct TileGeom { int workers, ctas; };

template <int K>
int launch_k(const EncodeParams &p, const EncodeMode &md, uint32_t max_wave_len, cudaStream_t st)
{
    const int g_num_sms = device_sm_count();
    const size_t smem_multi = (size_t)(kEncMaxThreads * 13 + 1 + 2 * kEncMaxThreads + 33 + 3) * sizeof(uint32_t);
    if (max_wave_len > (uint32_t)kEncTileMaxL) {
        // few long waves: several CTAs per wave (three launches); many: one CTA per wave
        const uint32_t maxseg = encode_long_maxseg(max_wave_len);
        if (md.long_scratch && p.nwaves < 4u * (uint32_t)g_num_sms && p.nwaves <= 1024u &&
            encode_long_scratch_bytes(p.nwaves, max_wave_len) <= md.long_scratch_bytes) {
            LongParams lp;
            lp.p = p;
            lp.seg_bits = (uint64_t *)md.long_scratch;
            lp.wave_off = lp.seg_bits + (size_t)p.nwaves * maxseg;
            lp.maxseg = maxseg;
            lp.dmask = md.delta ? -1 : 0;
            static DeviceOnce once;
            if (once.first())
                cudaFuncSetAttribute(encode_long_pack_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_multi);
            const dim3 grid(maxseg - 1, p.nwaves);               // (the row's last entry is the wave's end, not a segment)
            encode_long_size_kernel<K><<<grid, kEncMaxThreads, 0, st>>>(lp);
            encode_long_scan_kernel<<<1, 1024, 0, st>>>(lp);
            encode_long_pack_kernel<K><<<grid, kEncMaxThreads, smem_multi, st>>>(lp);
            return 3;
        }
        encode_multi_kernel<K><<<p.nwaves, kEncMaxThreads, smem_multi, st>>>(p, md.delta ? -1 : 0);
        return 1;
    }
    static const int lut_env = [] { const char *v = getenv("DRICE_ENC_LUT"); return v ? atoi(v) : 1; }();
    static const int workers_env = [] { const char *v = getenv("DRICE_ENC_WORKERS"); return v ? atoi(v) : 0; }();
    static const long stage_env = [] { const char *v = getenv("DRICE_ENC_STAGE_WORDS"); return v ? atol(v) : 0l; }();
    bool lut = LutConst<K>::kOk && md.delta && lut_env != 0;
    size_t table = lut ? lut_table_bytes<K, 1>() : 0;
    const uint32_t worst = (25u * max_wave_len + 31u) / 32u + 24u;
    uint32_t want = md.words_hint ? md.words_hint + md.words_hint / 16u + 24u : (10u * max_wave_len + 31u) / 32u + 24u;
    if (stage_env > 0) want = (uint32_t)stage_env;
    if (want > worst) want = worst;
    if (want < 64u) want = 64u;
    // room of a geometry: (227 KB - 1 KB reserved per CTA) / CTAs - static shared memory - table
    auto room_words = [&](TileGeom g) -> uint32_t {
        const size_t per_cta = (size_t)(227 * 1024) / g.ctas - 1024 - (160 + 144 * (size_t)g.workers) - table;   // (static: rings of tile state + wave slots)
        return (uint32_t)(per_cta / ((size_t)g.workers * 8)) & ~3u;
    };
    auto pick = [&]() -> TileGeom {
        // small batches (up to two rounds of 2 x 8-wave tiles per SM): tiles of 8 waves spread over more SMs and
        // finish sooner (2000 x 7000: 0.045 against 0.052 ms; 20 x 7000: 0.038 against 0.045)
        const bool small = !workers_env && p.nwaves <= 32u * (uint32_t)g_num_sms;
        const TileGeom order_big[] = {{24, 1}, {12, 2}, {8, 2}, {8, 1}}, order_small[] = {{8, 2}, {8, 1}, {8, 1}, {8, 1}};
        const TileGeom (&order)[4] = small ? order_small : order_big;
        TileGeom geom = small ? order[1] : order[2];
        bool found = false;
        for (const TileGeom g : order) {
            if (!md.delta && g.workers == 24) continue;                // (no 24-worker kernel without the delta)
            if (workers_env && g.workers != workers_env) continue;
            if (room_words(g) >= want || (!md.words_hint && stage_env <= 0 && g.workers == (small ? 8 : md.delta ? 24 : 12))) { geom = g; found = true; break; }
        }
        if (!found && workers_env) for (const TileGeom g : order) if (g.workers == workers_env) { geom = g; break; }
        return geom;
    };
    TileGeom geom = pick();
    if (lut && !workers_env) {
        // the table takes room from the staging: when that costs resident worker warps (long records, e.g.
        // L = 7000 at RiceParameter 64: 1 x 8 instead of 2 x 8), the arithmetic front-end is the faster one
        const size_t t = table;
        lut = false; table = 0;
        const TileGeom ga = pick();
        if (ga.workers * ga.ctas > geom.workers * geom.ctas) geom = ga;
        else { lut = true; table = t; }
    }
    uint32_t stage = room_words(geom) < want ? room_words(geom) : want;
    if (!md.words_hint && stage_env <= 0) stage = room_words(geom) < worst ? room_words(geom) : worst;   // no hint: all the room
    stage = (stage + 3u) & ~3u;
    if (stage > room_words(geom)) stage = room_words(geom);
    {
        static const bool debug = getenv("DRICE_DEBUG") != nullptr;
        if (debug) fprintf(stderr, "[drice] encode_tile K=%d L=%u waves=%u: %d workers x %d CTAs, staging %u words (want %u, hint %u), %s front-end\n",
                           K, max_wave_len, p.nwaves, geom.workers, geom.ctas, stage, want, md.words_hint, lut ? "table" : "arithmetic");
    }
    auto launch = [&](auto kernel, DeviceOnce &attr_set, int nworkers) {
        const size_t smem = (size_t)stage * 2 * nworkers * sizeof(uint32_t) + table;   // two buffers per worker warp
        const int nthreads = (nworkers + 1) * 32;
        const uint32_t ntiles = (p.nwaves + nworkers - 1) / nworkers;
        if (attr_set.first()) {
            cudaFuncAttributes fa;                              // (static + dynamic <= 227 KB)
            cudaFuncGetAttributes(&fa, kernel);
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes);
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        }
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, nthreads, smem);
        if (occ < 1) occ = 1;
        uint32_t grid = (uint32_t)(g_num_sms * occ);
        if (grid > ntiles) grid = ntiles;
        kernel<<<grid, nthreads, smem, st>>>(p, stage, ntiles, md.max_words);
    };
    static DeviceOnce a12, a24, a8, a8n, t12, t24, t8;                // per K (this function is a template)
    if constexpr (LutConst<K>::kOk) {
        if (lut) {
            if (geom.workers == 12) launch(encode_tile_kernel<K, 2, true, 12, 1>, t12, 12);
            else if (geom.workers == 24) launch(encode_tile_kernel<K, 1, true, 24, 1>, t24, 24);
            else launch(encode_tile_kernel<K, 2, true, 8, 1>, t8, 8);
            return 1;
        }
    }
    if (!md.delta) launch(encode_tile_kernel<K, 2, false, 8>, a8n, 8);    // no delta (filter [1] / pre-filtered input)
    else if (geom.workers == 24) launch(encode_tile_kernel<K, 1, true, 24>, a24, 24);
    else if (geom.workers == 12) launch(encode_tile_kernel<K, 2, true, 12>, a12, 12);
    else launch(encode_tile_kernel<K, 2, true, 8>, a8, 8);
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
// rows of the long-wave encoder's segment table: segments of the longest wave (+ 1 slot of misalignment) + the end entry
uint32_t encode_long_maxseg(uint32_t max_wave_len)
{
    const uint64_t slots = ((uint64_t)max_wave_len + 15u + 15u) / 16u;
    return (uint32_t)((slots + 1024u - 1u) / 1024u) + 1u;
}
size_t encode_long_scratch_bytes(uint32_t nwaves, uint32_t max_wave_len)
{
    if (max_wave_len <= (uint32_t)kEncTileMaxL) return 0;
    return ((size_t)nwaves * encode_long_maxseg(max_wave_len) + nwaves) * sizeof(uint64_t);
}
#endif
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
