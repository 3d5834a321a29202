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

#include <cstdlib>

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
// a * b + c with a 64-bit result: IMAD.WIDE.U32 (FMA pipe).  With b = 2^len this is
// "shift a left by len inside a 64-bit window and append c".
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c)
{
    uint64_t d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
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
template <bool kGuard>
struct Packer {
    uint32_t lo, n;
    uint32_t *ptr;
    uint32_t *end;      // kGuard: nothing is stored at or past `end` (packing straight into HBM)

    __device__ __forceinline__ void store(uint32_t *q, uint32_t v) const
    {
        if (!kGuard || q < end) *q = v;
    }

    // append one code of len <= 31 bits
    __device__ __forceinline__ void put(uint32_t v, uint32_t len)
    {
        const uint64_t a = mad_wide(lo, pow2(len), (uint64_t)v);
        n += len;
        if (n >= 32u) {
            n -= 32u;
            store(ptr++, __funnelshift_r((uint32_t)a, (uint32_t)(a >> 32), n));
        }
        lo = (uint32_t)a;
    }
    // append two codes of len <= 31 bits each with one flush of up to two words
    __device__ __forceinline__ void put2(uint32_t vA, uint32_t lenA, uint32_t vB, uint32_t lenB)
    {
        const uint32_t eB = pow2(lenB);
        const uint64_t a = mad_wide(lo, pow2(lenA), (uint64_t)vA);
        const uint64_t r10 = mad_wide((uint32_t)a, eB, (uint64_t)vB);
        const uint64_t r21 = mad_wide((uint32_t)(a >> 32), eB, r10 >> 32);
        const uint32_t r0 = (uint32_t)r10, r1 = (uint32_t)r21, r2 = (uint32_t)(r21 >> 32);
        n += lenA + lenB;
        const uint32_t k = n >> 5;                       // 0, 1 or 2 finished words
        const uint32_t X = __funnelshift_r(r1, r2, n), Y = __funnelshift_r(r0, r1, n);
        if (k == 2u) store(ptr, X);
        ptr += k;
        if (k != 0u) store(ptr - 1, Y);
        n &= 31u;
        lo = r0;
    }
};

// 16 consecutive samples of one lane as 8 packed words.  `q` = address of the lane's first
// sample; the widest naturally aligned vector load the wave's start allows is used (warp
// uniform `mis` = (address of the wave's first sample mod 16) / 2).  `edge`: the slot may
// leave the batch buffer [.., hi): read element-wise.
struct RawWords { uint32_t w[8]; };

__device__ __forceinline__ RawWords load_round(const int16_t *q, uint32_t mis, bool active, bool edge,
                                               const int16_t *hi)
{
    RawWords r;
#pragma unroll
    for (int m = 0; m < 8; ++m) r.w[m] = 0;
    if (!active) return r;
    if (edge && q + S > hi) {
#pragma unroll
        for (int i = 0; i < S; ++i)
            if (q + i < hi) r.w[i >> 1] |= (uint32_t)(uint16_t)q[i] << (16 * (i & 1));
        return r;
    }
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
    return r;
}

// exclusive word offset of tile g among all tiles: decoupled look-back by one warp, four
// rows of 32 status words in flight per round trip
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
                    __nanosleep(100);
                    if ((v >> 62) == 0) v = ld_relaxed_u64(lookback + my);
                }
            }
            idx -= 128;
        }
        if (lane == 0) st_relaxed_u64(lookback + g, kFlagPrefix | (excl + mine));
    }
    return excl;
}

// per-wave addressing
struct WaveView {
    const int16_t *wave, *raw_hi;
    uint32_t mis;               // (address of the first sample mod 16) / 2
    uint32_t n, nrounds;
    bool edge_hi;               // slots of the last round can leave the batch buffer

    __device__ __forceinline__ WaveView(const EncodeParams &p, const WaveGeom &wg)
    {
        wave = p.raw + wg.begin;
        raw_hi = p.raw + p.raw_samples;
        mis = (uint32_t)((reinterpret_cast<uintptr_t>(wave) & 15u) >> 1);
        n = wg.n;
        nrounds = (n + kRound - 1) / kRound;
        edge_hi = wave + (size_t)nrounds * kRound > raw_hi;
    }
    __device__ __forceinline__ RawWords load(uint32_t r, int lane) const
    {
        const uint32_t s0 = r * kRound + lane * S;
        return load_round(wave + s0, mis, s0 < n, r + 1 == nrounds && edge_hi, raw_hi);
    }
};

// packed words of one lane and round -> packed zig-zag values; `prev_last` carries the last
// word of the previous round's lane 31 (0 at the start of a wave: d[0] = x[0], :53-56)
__device__ __forceinline__ void delta_zigzag(uint32_t (&w)[8], uint32_t nvalid, uint32_t &prev_last, int lane,
                                             uint32_t (&U)[8], uint32_t &uor)
{
    // short last slot: repeat the last valid sample; its codes (delta 0) trail the lane's bits
    // and are cut off by the callers
    if (nvalid > 0 && nvalid < (uint32_t)S) {
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            if (2u * v + 1 == nvalid) w[v] = prmt(w[v], 0, 0x1010);
            if (v > 0 && 2u * v >= nvalid) w[v] = prmt(w[v - 1], 0, 0x3232);
        }
    }
    // word holding the sample before this lane's first one in its HIGH half
    uint32_t pw = __shfl_up_sync(0xffffffffu, w[7], 1);
    if (lane == 0) pw = prev_last;
    prev_last = __shfl_sync(0xffffffffu, w[7], 31);
    // D = per-half (x[j] - x[j-1]);  U = (D + D) ^ sign(D)   (src/deltaRice.c:57-62, :207-211)
    uor = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const uint32_t prev = m ? w[m - 1] : pw;
        const uint32_t X = w[m] * 0xFFFF0001u;              // high half: hi(w) - lo(w)
        const uint32_t Y = w[m] - (prev >> 16);             // low half:  lo(w) - hi(prev)
        const uint32_t D = prmt(Y, X, 0x7610);
        const uint32_t Sg = prmt(D, 0, 0xbb99);             // per-half sign mask
        U[m] = __vadd2(D, D) ^ Sg;
        uor |= U[m];
    }
}

template <int K>
struct RiceConst {
    static constexpr bool kPairs = (K >= 1 && K <= 7);   // two samples merge into one code of <= 30 bits
    static constexpr uint32_t M = 1u << K;
    static constexpr uint32_t MM = M * 0x10001u, NN = (2u * M - 1u) * 0x10001u, QM = 0xFFFFu >> K;
    static constexpr uint32_t HM = ((0xFFFFu << ((K + 3) > 16 ? 16 : (K + 3))) & 0xFFFFu) * 0x10001u;   // quotient >= 8
};

// encoding sweep.  kDirect = false: packs into the warp's staging of `cap` words; when the wave
// outgrows it, packing stops (sizing continues) and *overflow is set.  kDirect = true: packs
// straight into the record in HBM, `cap` = the wave's word count (nothing is stored past it).
// Returns the wave's bit count.
template <int K, bool kDirect>
__device__ __forceinline__ uint32_t encode_wave(const WaveView &wv, int lane, uint32_t *dst, uint32_t cap, bool *overflow)
{
    bool ovf = false;
    using C = RiceConst<K>;
    constexpr bool kPairs = C::kPairs;
    constexpr uint32_t M = C::M;
    uint32_t base = 0;                  // bits packed so far
    uint32_t carry_round = 0;           // pending bits (left aligned) of the previous round's last lane
    uint32_t prev_last = 0;
    RawWords cur = wv.load(0, lane);
    for (uint32_t r = 0; r < wv.nrounds; ++r) {
        const uint32_t s0 = r * kRound + lane * S;
        const int32_t rem = (int32_t)wv.n - (int32_t)s0;
        const uint32_t nvalid = rem >= S ? (uint32_t)S : (rem > 0 ? (uint32_t)rem : 0u);
        RawWords nxt = cur;
        if (r + 1 < wv.nrounds) nxt = wv.load(r + 1, lane);
        uint32_t U[8], uor;
        delta_zigzag(cur.w, nvalid, prev_last, lane, U, uor);

        // ---- Rice codes: items (value, length) kept in registers across the scan ---------
        // kPairs: item m = samples 2m, 2m+1 merged into one code of <= 30 bits; an item that
        // holds an escape is flagged (bit 7 of its length) and keeps the packed zig-zag
        // values instead.  !kPairs: 16 single-sample items.
        constexpr int NI = kPairs ? 8 : S;
        uint32_t iv[NI], il[NI];
        uint32_t T = 0;                     // bits of this lane
        bool flagged = false;
        if (kPairs) {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const uint32_t u2 = U[m];
                const uint32_t V2 = (u2 | C::MM) & C::NN;
                const uint32_t qhi = u2 >> (16 + K), qlo = (u2 >> K) & C::QM;
                il[m] = qlo + qhi + 2u * (K + 1);
                iv[m] = (V2 & 0xFFFFu) * ((2u * M) << qhi) + (V2 >> 16);
                T += il[m];
            }
            flagged = (uor & C::HM) != 0u;
            if (flagged) {                  // rare and divergent: redo the items that hold an escape
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
            // padding samples of a short last slot were coded as delta 0: K+1 bits each
            if (nvalid < (uint32_t)S) T = nvalid ? T - ((uint32_t)S - nvalid) * (K + 1) : 0u;
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const uint32_t u = (j & 1) ? (U[j >> 1] >> 16) : (U[j >> 1] & 0xFFFFu);
                rice_code<K>(u, iv[j], il[j]);
                if ((uint32_t)j >= nvalid) { iv[j] = 0; il[j] = 0; }
                T += il[j];
            }
        }

        // ---- warp exclusive scan of T -----------------------------------------------------
        uint32_t inc = T;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        const uint32_t b0 = base + inc - T;            // bit offset of this lane in the wave

        // ---- pack ------------------------------------------------------------------------
        Packer<kDirect> pk;
        pk.n = b0 & 31u;
        pk.ptr = dst + (b0 >> 5);
        pk.end = dst + cap;
        pk.lo = 0;
        if (!kDirect && ((base + total + 31u) >> 5) + 16u > cap) ovf = true;   // warp uniform
        const bool packs = nvalid > 0 && !ovf;
        if (packs) {
            if (kPairs) {
#pragma unroll
                for (int m = 0; m < 8; m += 2) {
                    if (flagged && ((il[m] | il[m + 1]) & 0x80u)) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (il[m + e] & 0x80u) {
                                uint32_t v0, l0, v1, l1;
                                rice_code<K>(iv[m + e] & 0xFFFFu, v0, l0);
                                rice_code<K>(iv[m + e] >> 16, v1, l1);
                                pk.put(v0, l0);
                                pk.put(v1, l1);
                            } else {
                                pk.put(iv[m + e], il[m + e]);
                            }
                        }
                    } else {
                        pk.put2(iv[m], il[m], iv[m + 1], il[m + 1]);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < S; ++j)
                    if (il[j]) pk.put(iv[j], il[j]);
            }
        }
        // ---- stitch the lanes: the bits a lane left pending belong to the first word the
        // next lane wrote (or still holds) -------------------------------------------------
        uint32_t frag;                                       // pending bits, left aligned
        asm("shl.b32 %0, %1, %2;" : "=r"(frag) : "r"(pk.lo), "r"(32u - pk.n));   // n == 0 -> 0
        if (nvalid == 0) frag = 0;
        const bool flushed = pk.ptr != dst + (b0 >> 5);      // wrote its first word itself
        constexpr int kStitch = (K == 0) ? 2 : 1;            // 1-bit codes: a lane may hold < 32 bits
#pragma unroll
        for (int e = 0; e < kStitch; ++e) {
            uint32_t from_prev = __shfl_up_sync(0xffffffffu, frag, 1);
            if (lane == 0) from_prev = carry_round;
            if (packs) {
                if (flushed) { if (from_prev) dst[b0 >> 5] |= from_prev; }
                else frag |= from_prev;
            }
        }
        carry_round = __shfl_sync(0xffffffffu, frag, 31);
        // the wave's last lane owns the final partial word; bits past the wave's end (padding
        // codes of a short slot) are cleared so the word is zero padded (:237-241)
        base += total;
        if (r + 1 == wv.nrounds && packs && s0 + S >= wv.n) {
            if (pk.n) pk.store(pk.ptr, frag);
            if (base & 31u) dst[base >> 5] &= 0xFFFFFFFFu << (32u - (base & 31u));
        }
        cur = nxt;
    }
    __syncwarp();
    *overflow = ovf;
    return base;
}

// ---- tile kernel ----------------------------------------------------------------------------
// A tile = kEncWarps consecutive waves, taken by one persistent CTA of kEncWarps worker warps +
// one control warp.  Per iteration a worker encodes ONE wave of the current tile into one of its
// two staging buffers (single sweep over HBM), then copies out the wave it encoded in the
// previous iteration, whose position has been resolved in the meantime by the control warp:
// after the one barrier of the iteration the control warp sums the tile's wave sizes, publishes
// the aggregate and resolves the tile's offset among all tiles by decoupled look-back, while the
// workers are already encoding the next tile.  No worker waits on global memory latency, and
// only one look-back per tile is in flight per CTA.
template <int K>
__global__ void __launch_bounds__((kEncWarps + 1) * 32)
encode_tile_kernel(const EncodeParams p, const uint32_t stage_words, const uint32_t ntiles)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t s_tile[2];
    __shared__ uint32_t s_mine[2][kEncWarps];
    __shared__ uint64_t s_off[2];
    __shared__ volatile uint32_t s_flag[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool control = warp == kEncWarps;

    if (threadIdx.x == 0) { s_flag[0] = 0; s_flag[1] = 0; }
    if (control && lane == 0) s_tile[0] = atomicAdd(p.ticket, 1u);
    __syncthreads();

    if (control) {
        for (uint32_t it = 0;; ++it) {
            const int par = it & 1;
            const uint32_t tile = s_tile[par];
            if (tile >= ntiles) break;
            if (lane == 0) s_tile[par ^ 1] = atomicAdd(p.ticket, 1u);
            __syncthreads();                                 // B(it): wave sizes of the tile
            const uint32_t v = lane < kEncWarps ? s_mine[par][lane] : 0u;
            const uint64_t mine = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) st_relaxed_u64(p.lookback + tile, (tile == 0 ? kFlagPrefix : kFlagAggregate) | mine);
            const uint64_t excl = lookback_excl(p.lookback, tile, mine, lane);
            if (lane == 0) {
                s_off[par] = excl;
                __threadfence_block();
                s_flag[par] = it + 1;
                if (excl + mine > p.out_cap_words) atomicOr(p.status, kErrCapacity);
                if (tile == ntiles - 1) p.chunk_byte_off[p.nchunks] = (excl + mine) * 4;
            }
            __syncwarp();
        }
        return;
    }

    // ---- workers ---------------------------------------------------------------------------
    uint32_t *const stage0 = smem + (size_t)(2 * warp) * stage_words;   // two staging buffers per warp
    WaveGeom wg_prev;
    uint32_t nwords_prev = 0;
    bool have_prev = false, ovf_prev = false;
    wg_prev.g = 0xffffffffu;
    for (uint32_t it = 0;; ++it) {
        const int par = it & 1;
        const uint32_t tile = s_tile[par];
        const bool live = tile < ntiles;
        WaveGeom wg;
        uint32_t nwords = 0;
        bool have = false, ovf = false;
        if (live) {
            const uint32_t g = tile * kEncWarps + warp;
            uint32_t mine = 0;
            if (g < p.nwaves) {
                have = true;
                wg = locate_wave(p, g);
                if (wg.chunk_total) {
                    const WaveView wv(p, wg);
                    nwords = (encode_wave<K, false>(wv, lane, stage0 + par * stage_words, stage_words, &ovf) + 31u) >> 5;
                    mine = nwords + 1u;
                }
                mine += wg.first;                            // empty chunk: header only
            }
            if (lane == 0) s_mine[par][warp] = mine;
        }
        // ---- copy out the wave of the previous iteration ------------------------------------
        if (have_prev) {
            const int pp = par ^ 1;
            const uint32_t v = lane < kEncWarps ? s_mine[pp][lane] : 0u;
            const uint32_t loff = __reduce_add_sync(0xffffffffu, lane < warp ? v : 0u);
            while (s_flag[pp] != it) __nanosleep(40);        // tile offset: normally there long ago
            __threadfence_block();
            const uint64_t off = s_off[pp] + loff;
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
                        const uint32_t *src = stage0 + pp * stage_words;
                        for (uint32_t i = lane; i < nwords_prev; i += 32) rec[1 + i] = src[i];
                    } else {                                 // larger than the staging: pack in place
                        const WaveView wv(p, wg_prev);
                        bool dummy;
                        encode_wave<K, true>(wv, lane, rec + 1, nwords_prev, &dummy);
                    }
                }
            }
            __syncwarp();
        }
        if (!live) break;
        wg_prev = wg;
        nwords_prev = nwords;
        have_prev = have;
        ovf_prev = ovf;
        __syncthreads();                                     // B(it)
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

// zig-zag of the 16-bit wrapped difference cur-prev (src/deltaRice.c:57-62, :207-211)
__device__ __forceinline__ uint32_t zigzag_delta(int cur, int prev)
{
    const int t = cur - prev;
    const uint32_t s = (uint32_t)((int)((uint32_t)t << 16) >> 31);
    return (((uint32_t)t << 1) ^ s) & 0xFFFFu;
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
                                               uint32_t (&cv)[kSamplesPerThread])
{
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
            cv[j] = rice_code_packed<K>(zigzag_delta(x[j + 1], x[j]));
            T += cv[j] >> 24;
        }
    } else {
        x[0] = (jlo == 0 && s0 > 0) ? (int)sp[-1] : 0;
#pragma unroll
        for (int j = 0; j < S; ++j) x[j + 1] = (j >= jlo && j < jhi) ? (int)sp[j] : 0;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const uint32_t c = rice_code_packed<K>(zigzag_delta(x[j + 1], x[j]));
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
__device__ __forceinline__ void pack_wave_streaming(const int16_t *wave, uint32_t n, uint32_t *rec, uint32_t *smem)
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
        const uint32_t T = slot_codes<K>(wave, ((int64_t)t * NT + tid) * S - mis, n, cv);
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
encode_multi_kernel(const EncodeParams p)
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
        bits += slot_codes<K>(wave, ((int64_t)t * NT + tid) * S - mis, wg.n, cv);
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
    pack_wave_streaming<K>(wave, wg.n, rec, smem);
}

int g_num_sms = 0;

uint32_t stage_cap_words()
{
    static uint32_t v = 0;
    if (!v) {
        const char *e = getenv("DRICE_ENC_STAGE_WORDS");
        long w = e ? atol(e) : 1600;
        if (w < 64) w = 64;
        if (w > 12000) w = 12000;
        v = (uint32_t)w;
    }
    return v;
}

template <int K>
int launch_k(const EncodeParams &p, uint32_t max_wave_len, cudaStream_t st)
{
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    const size_t smem_multi = (size_t)(kEncMaxThreads * 13 + 1 + 2 * kEncMaxThreads + 33 + 3) * sizeof(uint32_t);
    if (max_wave_len > (uint32_t)kEncTileMaxL) {
        encode_multi_kernel<K><<<p.nwaves, kEncMaxThreads, smem_multi, st>>>(p);
        return 1;
    }
    // per-warp staging: the worst case of the longest wave when it is affordable, else a cap
    // (a wave that outgrows it is packed straight into its record in HBM)
    uint32_t stage = (25u * max_wave_len + 31u) / 32u + 24u;
    if (stage > stage_cap_words()) stage = stage_cap_words();
    stage = (stage + 3u) & ~3u;
    const size_t smem = (size_t)stage * 2 * kEncWarps * sizeof(uint32_t);   // two buffers per worker warp
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(encode_tile_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(encode_tile_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr_set = true;
    }
    const int nthreads = (kEncWarps + 1) * 32;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, encode_tile_kernel<K>, nthreads, smem);
    if (occ < 1) occ = 1;
    const uint32_t ntiles = (p.nwaves + kEncWarps - 1) / kEncWarps;
    uint32_t grid = (uint32_t)(g_num_sms * occ);
    if (grid > ntiles) grid = ntiles;
    encode_tile_kernel<K><<<grid, nthreads, smem, st>>>(p, stage, ntiles);
    return 1;
}

}  // namespace

int launch_encode(const EncodeParams &p, uint32_t max_wave_len, cudaStream_t st)
{
    if (p.nwaves == 0) return 0;
    switch (p.k) {
#define DRICE_CASE(K) case K: return launch_k<K>(p, max_wave_len, st);
        DRICE_CASE(0) DRICE_CASE(1) DRICE_CASE(2) DRICE_CASE(3) DRICE_CASE(4) DRICE_CASE(5)
        DRICE_CASE(6) DRICE_CASE(7) DRICE_CASE(8) DRICE_CASE(9) DRICE_CASE(10) DRICE_CASE(11)
        DRICE_CASE(12) DRICE_CASE(13) DRICE_CASE(14) DRICE_CASE(15)
#undef DRICE_CASE
    }
    return -1;
}

}  // namespace drice
