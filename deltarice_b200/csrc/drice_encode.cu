// drice_encode.cu — Delta + Rice ENCODE for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   encodeWaveform delta branch        src/deltaRice.c:49-63
//   compressWithRiceCoding             src/deltaRice.c:191-244
//   perWaveCompression                 src/deltaRice.c:365-381
//   writeWholeCompressedByteString     src/deltaRice.c:383-436 (framing + compaction)
//
// One CTA per wave ("waveform" of L samples), single pass over HBM:
//   1. every thread owns one 32-byte ALIGNED slot of 16 samples (two 128-bit loads; the
//      slot grid is aligned to memory, not to the wave, so waves may start anywhere),
//      computes delta -> zig-zag -> (value,length) per sample and its bit total;
//   2. block exclusive scan of the bit totals (warp shuffles + one smem hop) gives every
//      thread its bit offset and the wave its word count, which is published at once for
//      the cross-wave scan (decoupled look-back over a 64-bit status word per wave,
//      waves take tickets so a CTA only ever waits on CTAs that already started);
//   3. threads pack their codes MSB-first into 32-bit words in shared memory.  Words are
//      written by the thread that STARTS them; the leading fragment a thread contributes
//      to a word started by a predecessor goes to a side array and is OR-ed in by that
//      word's owner after one barrier - no shared-memory atomics;
//   4. the record [nwords][words] is copied out coalesced at the scanned offset; the first
//      wave of a chunk also writes the chunk header [total].
// Waves longer than one CTA tile (L > 8177) run the same code in a tile loop
// (kMulti = true): a sizing sweep, then a packing sweep that re-reads the wave (L1/L2
// resident for moderate L) and streams completed words straight to HBM.
#include "drice_kernels.cuh"

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
__device__ __forceinline__ int4 ld_stream_v4(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// zig-zag of the 16-bit wrapped difference cur-prev (src/deltaRice.c:57-62, :207-211):
// d = (int16)(cur-prev); u = d >= 0 ? 2d : -2d-1  ==  ((t<<1) ^ -(bit15 of t)) & 0xffff
__device__ __forceinline__ uint32_t zigzag_delta(int cur, int prev)
{
    const int t = cur - prev;
    const uint32_t s = (uint32_t)((int)((uint32_t)t << 16) >> 31);
    return (((uint32_t)t << 1) ^ s) & 0xFFFFu;
}

// (value, length) of one sample packed as value | length << 24 (src/deltaRice.c:212-228)
template <int K>
__device__ __forceinline__ uint32_t rice_code(uint32_t u)
{
    constexpr uint32_t M = 1u << K;
    const uint32_t q = u >> K;
    uint32_t len = q + (K + 1);
    uint32_t val = (u & (M - 1u)) | M;
    if (q >= kEscapeQuotient) {
        len = kEscapeBits;
        val = u | 0x10000u;
    }
    return val | (len << 24);
}

struct WaveGeom {
    uint64_t begin;     // first sample of the wave in raw
    uint32_t n;         // samples in the wave
    uint32_t chunk;     // chunk index
    uint32_t first;     // 1 if first wave of its chunk
    uint64_t chunk_total;
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
    w.chunk_total = ce - cb;
    return w;
}

// Codes of the 16-sample slot whose first sample has wave-relative index s0 (may be negative
// or run past n: such samples get length 0).  Returns the slot's bit total.
template <int K>
__device__ __forceinline__ uint32_t slot_codes(const int16_t *wave, int64_t s0, uint32_t n,
                                               uint32_t (&cv)[kSamplesPerThread])
{
    constexpr int S = kSamplesPerThread;
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
        // full slot: 32-byte aligned by construction
        const int4 *vp = reinterpret_cast<const int4 *>(sp);
        const int4 a = ld_stream_v4(vp), b = ld_stream_v4(vp + 1);
        const int w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        x[0] = (s0 > 0) ? (int)sp[-1] : 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            x[2 * m + 1] = (int)(short)(w[m] & 0xFFFF);
            x[2 * m + 2] = w[m] >> 16;
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
            cv[j] = rice_code<K>(zigzag_delta(x[j + 1], x[j]));
            T += cv[j] >> 24;
        }
    } else {
        x[0] = (jlo == 0 && s0 > 0) ? (int)sp[-1] : 0;
#pragma unroll
        for (int j = 0; j < S; ++j) x[j + 1] = (j >= jlo && j < jhi) ? (int)sp[j] : 0;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const uint32_t c = rice_code<K>(zigzag_delta(x[j + 1], x[j]));
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

// Packs the slot's codes into shared words starting at tile-local bit position b0.
// Word ownership: see file header.  Returns via refs the pending tail word.
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

template <int K, bool kMulti>
__global__ void __launch_bounds__(kEncMaxThreads)
encode_kernel(const EncodeParams p)
{
    constexpr int S = kSamplesPerThread;
    extern __shared__ __align__(16) uint32_t smem[];
    const int NT = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *sbits = smem;                  // NT*13 + 1 (worst case 12.5 words per slot)
    uint32_t *shead = sbits + NT * 13 + 1;   // NT
    uint32_t *sboff = shead + NT;            // NT
    uint32_t *swarp = sboff + NT;            // 33
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
    const uint32_t ntiles = kMulti ? (nslots + NT - 1) / NT : 1u;

    uint32_t cv[S];
    uint32_t T = 0, boff = 0, nwords;

    // ---- sizing ------------------------------------------------------------------
    if (!kMulti) {
        T = slot_codes<K>(wave, (int64_t)tid * S - mis, wg.n, cv);
        uint32_t total;
        boff = block_excl_scan(T, swarp, &total);
        nwords = (total + 31u) >> 5;
    } else {
        uint64_t bits = 0;
        for (uint32_t t = 0; t < ntiles; ++t)
            bits += slot_codes<K>(wave, ((int64_t)t * NT + tid) * S - mis, wg.n, cv);
        // block reduce (64-bit via two 32-bit scans would overflow; use shuffles + smem)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, d);
        uint64_t *s64 = reinterpret_cast<uint64_t *>(sbits);
        if (lane == 0) s64[warp] = bits;
        __syncthreads();
        uint64_t tot = 0;
        for (int w = 0; w < (NT >> 5); ++w) tot += s64[w];
        __syncthreads();
        nwords = (uint32_t)((tot + 31u) >> 5);
    }
    const uint32_t rec_words = wg.chunk_total ? nwords + 1u : 0u;   // empty chunk: header only
    const uint64_t mine = (uint64_t)rec_words + wg.first;

    // ---- publish this wave's size for the cross-wave scan ----------------------------
    if (tid == 0) st_relaxed_u64(p.lookback + g, (g == 0 ? kFlagPrefix : kFlagAggregate) | mine);

    // ---- single tile: pack into shared memory while predecessors publish --------------
    bool owner = false;
    uint32_t tail = 0, wt = 0;
    if (!kMulti) {
        shead[tid] = 0;
        sboff[tid] = boff;
        pack_slot(cv, boff, T, sbits, shead, owner, tail, wt);
        __syncthreads();
        if (owner) {
            for (uint32_t j = tid + 1; j < nslots && (sboff[j] >> 5) == wt; ++j) tail |= shead[j];
            sbits[wt] = tail;
        }
    }

    // ---- decoupled look-back (warp 0) -------------------------------------------------
    if (warp == 0) {
        uint64_t excl = 0;
        if (g > 0) {
            int64_t idx = (int64_t)g - 1;
            while (true) {
                const int64_t my = idx - lane;
                uint64_t s = kFlagPrefix;
                if (my >= 0) {
                    s = ld_relaxed_u64(p.lookback + my);
                    while ((s >> 62) == 0) {
                        __nanosleep(32);
                        s = ld_relaxed_u64(p.lookback + my);
                    }
                }
                const uint32_t pm = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                uint64_t v = s & kValueMask;
                if (pm) {
                    const int firstp = __ffs(pm) - 1;
                    if (lane > firstp) v = 0;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                excl += v;
                if (pm) break;
                idx -= 32;
            }
            if (lane == 0) st_relaxed_u64(p.lookback + g, kFlagPrefix | (excl + mine));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const uint64_t off = s_excl;              // word offset of this wave's contribution
    const bool fits = off + mine <= p.out_cap_words;
    if (tid == 0) {
        if (!fits) atomicOr(p.status, kErrCapacity);
        if (wg.first) p.chunk_byte_off[wg.chunk] = off * 4;
        if (g == p.nwaves - 1) p.chunk_byte_off[p.nchunks] = (off + mine) * 4;
    }
    if (!fits) return;
    uint32_t *rec = p.out + off + wg.first;
    if (tid == 0) {
        if (wg.first) p.out[off] = (uint32_t)wg.chunk_total;
        if (rec_words) rec[0] = nwords;
    }
    if (rec_words == 0) return;

    if (!kMulti) {
        for (uint32_t w = tid; w < nwords; w += NT) rec[1 + w] = sbits[w];
        return;
    }

    // ---- multi tile: packing sweep, completed words stream to HBM ---------------------
    uint64_t P = 0;                            // bits emitted so far
    __shared__ uint32_t s_carry;
    if (tid == 0) s_carry = 0;
    for (uint32_t t = 0; t < ntiles; ++t) {
        T = slot_codes<K>(wave, ((int64_t)t * NT + tid) * S - mis, wg.n, cv);
        uint32_t total;
        boff = block_excl_scan(T, swarp, &total);
        const uint32_t pre = (uint32_t)(P & 31u);          // bits already in word 0 (carry)
        const uint32_t b0 = pre + boff;
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

template <int K>
int launch_k(const EncodeParams &p, uint32_t max_wave_len, cudaStream_t st)
{
    const bool multi = max_wave_len > (uint32_t)kEncTileMaxL;
    int nt;
    if (multi) {
        nt = kEncMaxThreads;
    } else {
        const uint32_t slots = (max_wave_len + 2 * kSamplesPerThread - 2) / kSamplesPerThread;
        nt = (int)((slots + 31) / 32) * 32;
        if (nt < 32) nt = 32;
        if (nt > kEncMaxThreads) nt = kEncMaxThreads;
    }
    const size_t smem = (size_t)(nt * 13 + 1 + nt + nt + 33 + 3) * sizeof(uint32_t);
    if (multi)
        encode_kernel<K, true><<<p.nwaves, nt, smem, st>>>(p);
    else
        encode_kernel<K, false><<<p.nwaves, nt, smem, st>>>(p);
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
