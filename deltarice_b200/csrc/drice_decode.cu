// drice_decode.cu — Rice DECODE + inverse delta for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   readWholeCompressedByteString   src/deltaRice.c:301-341 (header walk :319-325)
//   perWaveDecompression            src/deltaRice.c:293-297
//   decompressWithRiceCoding        src/deltaRice.c:138-189
//   decodeWaveform delta branch     src/deltaRice.c:78-90
//
// Locating the records of a chunk (the stream has no index, only the chain cur += word[cur] + 1):
//   scan_headers_kernel + rank_headers_kernel   waves of a few thousand samples: headers are FOUND
//                  by their value range on all SMs, ranked per chunk and verified against the chain;
//                  anything inconsistent falls back to the exact serial chase
//   locate_kernel  short waves / very long waves: one CTA per chunk streams the chunk through a
//                  128 KB circular buffer in shared memory (bulk async copies on mbarriers) while ONE
//                  thread chases the chain at shared-memory latency
//   chase_direct_kernel   very large batches: one warp per chunk walks the chain through global memory
//                  (every chunk at once: cheaper than reading tens of GB of stream to find the headers)
// Decoding the records:
//   parse_kernel   one LANE per wave, 32 waves per warp: Rice parsing is a serial chain per wave, so
//                  the parallelism is across waves.  One register of state per lane
//                  (sample << 16 | bit position), one table lookup and one add per code, several
//                  codes per 64-bit window, 16 samples per 32-byte store.  Batches with heavy records
//                  (escapes) take the HEAVY instance: waves sorted by bits per sample first
//                  (wave_hist_kernel / wave_scatter_kernel), and warps of heavy waves skip the table
//   parse_wide_kernel   batches too small to fill the machine with lanes: one CTA per wave,
//                  parallel INSIDE the wave (transition functions of four-word runs composed along
//                  the record, then the runs decode independently)
//   parse_long_kernel   few waves of many words (or a handful of short ones): several CTAs per wave,
//                  segments of 256 ... 6400 words chained along the record
#include "drice_kernels.cuh"

#include <cstdio>
#include <cstdlib>

namespace drice {

namespace {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// shared memory through 32-bit shared-window addresses (no generic address arithmetic in the loop)
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds32o(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------
// locate
// ------------------------------------------------------------------------------------
// The stream has no index, only the chain cur += word[cur] + 1 (src/deltaRice.c:319-325), and a
// hop through HBM costs a DRAM round trip.  One CTA per chunk therefore streams the chunk through
// a 128 KB circular buffer in shared memory while ONE thread chases the chain at shared-memory
// latency.  Three roles, no block barriers:
//   producer  (warp 1, one thread) issues the tiles as bulk async copies (cp.async.bulk: one
//             instruction per 32 KB tile, completion on the stage's `full` mbarrier) as soon as the
//             chaser has released the stage (`empty` mbarrier);
//   chaser    (warp 0, one thread) runs FOUR hops speculatively per iteration on byte offsets -
//             the dependent chain is one multiply-add, one mask and one LDS per hop, everything
//             else (the 64-bit record positions it stores, the bounds checks, handing tiles
//             back) hides behind it; hops that ran past the loaded data are discarded;
//   the other warps fill the part of the wave table that does not depend on the chain
//             (output position, sample count).
// Chunks whose waves are long (>= kLocDirectL samples: few, large records) are chased straight
// through global memory instead - streaming them would move far more than the hops need.
constexpr int      kLocThreads   = 128;
constexpr int      kLocTileWords = 8192;          // 32 KB per stage
constexpr int      kLocStages    = 4;
constexpr uint32_t kLocRingWords = kLocTileWords * kLocStages;
constexpr uint64_t kLocDirectL   = 32768;

// (mbarrier / bulk-copy wrappers: drice_kernels.cuh)

// issues the tile of words [A, A + kLocTileWords) into the stage at shared address `dst`
// ((comp + A) is 16-byte aligned; A may be negative by up to 3 words when comp itself is not):
// the (at most 3 + 3) words around the 16-byte aligned part inside [0, limit) with plain loads,
// then that part as one bulk copy that completes on `mbar`.
__device__ __forceinline__ void locate_issue_tile(uint32_t dst, uint32_t mbar, const uint32_t *comp, int64_t A, uint64_t limit)
{
    const int64_t lo = A < 0 ? A + 4 : A;
    int64_t hi = A + (int64_t)(((int64_t)limit - A) & ~3ll);          // same alignment class as A, <= limit
    if (hi > A + kLocTileWords) hi = A + kLocTileWords;
    if (hi < lo) hi = lo;
    for (int64_t w = A < 0 ? 0 : A; w < lo && (uint64_t)w < limit; ++w) sts32(dst + (uint32_t)(w - A) * 4u, comp[w]);
    for (int64_t w = hi; w < A + kLocTileWords && (uint64_t)w < limit; ++w) sts32(dst + (uint32_t)(w - A) * 4u, comp[w]);
    const uint32_t bytes = (uint32_t)(hi - lo) * 4u;
    if (bytes) {
        mbar_arrive_expect_tx(mbar, bytes);
        bulk_g2s(dst + (uint32_t)(lo - A) * 4u, comp + lo, bytes, mbar);
    } else {
        mbar_arrive(mbar);
    }
}

__global__ void __launch_bounds__(kLocThreads) locate_kernel(const LocateParams p)
{
    extern __shared__ __align__(128) uint32_t stile[];  // the circular buffer: kLocStages tiles
    __shared__ __align__(8) uint64_t s_full[kLocStages], s_empty[kLocStages];
    __shared__ uint32_t s_located;                       // waves the chase reached
    __shared__ volatile uint32_t s_done;                 // the chaser has finished (the producer may stop)
    const uint32_t c = blockIdx.x;
    const uint64_t wb = p.chunk_word_off[c], we = p.chunk_word_off[c + 1];
    const uint64_t sb = p.chunk_sample_off[c], se = p.chunk_sample_off[c + 1];
    const uint32_t g0 = p.chunk_wave_off[c];
    const uint32_t W = p.chunk_wave_off[c + 1] - g0;    // waves expected from the caller's sizes
    const uint64_t total = se - sb;
    const uint64_t Lw = p.L ? (uint64_t)p.L : total;

    if (we <= wb) {                                      // no stream at all
        if (threadIdx.x == 0) atomicOr(p.status, kErrStream);
        for (uint32_t w = threadIdx.x; w < W; w += kLocThreads) { p.wave_in[g0 + w] = wb; p.wave_out[g0 + w] = sb; p.wave_n[g0 + w] = 0; }
        return;
    }
    if (W == 0) {
        if (threadIdx.x == 0) {
            if (p.comp[wb] != (uint32_t)total) atomicOr(p.status, kErrTotal);
            if (we != wb + 1) atomicOr(p.status, kErrStream);
        }
        return;
    }
    const bool direct = Lw >= kLocDirectL || (we - wb) >= (1ull << 28);
    const uint32_t stile_s = (uint32_t)__cvta_generic_to_shared(stile);
    const uint32_t full_s = (uint32_t)__cvta_generic_to_shared(s_full);
    const uint32_t empty_s = (uint32_t)__cvta_generic_to_shared(s_empty);
    // misalignment of the global address: tiles start at word indices A with (comp + A) 16-byte aligned
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p.comp) >> 2) & 3u);
    const int64_t A0 = (int64_t)((wb + 1 + mis) & ~3ull) - (int64_t)mis;     // first word of tile 0
    const uint32_t end_rel = (uint32_t)((int64_t)we - A0);                    // chunk end, words from A0
    const uint32_t ntiles = (end_rel + kLocTileWords - 1) / kLocTileWords;
    if (threadIdx.x == 0) {
        s_done = 0;
        s_located = W;
        if (!direct) {
#pragma unroll
            for (int sidx = 0; sidx < kLocStages; ++sidx) {
                mbar_init(full_s + 8u * sidx, 1u);
                mbar_init(empty_s + 8u * sidx, 1u);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    if (threadIdx.x == 0 && direct) {
        // ---- long waves: hop through global memory --------------------------------------------
        if (p.comp[wb] != (uint32_t)total) atomicOr(p.status, kErrTotal);
        uint64_t cur = wb + 1;
        uint32_t w = 0;
        while (w < W && cur < we) {
            p.wave_in[g0 + w] = cur;
            cur += (uint64_t)__ldg(p.comp + cur) + 1ull;
            ++w;
        }
        if (w < W) s_located = w;
        if (w < W || cur != we) atomicOr(p.status, kErrStream);
    } else if (threadIdx.x == 0) {
        // ---- chaser -----------------------------------------------------------------------------
        // positions are BYTE offsets from tile 0 (32 bit: chunks of 1 GiB of stream and more go the
        // direct way); a ring address is (offset & mask) + the buffer's (uniform) base
        if (p.comp[wb] != (uint32_t)total) atomicOr(p.status, kErrTotal);
        const char *ring = reinterpret_cast<const char *>(stile);
        constexpr uint32_t kMask = kLocRingWords * 4u - 4u;
        constexpr uint32_t kTileBytes = kLocTileWords * 4u;
        const uint32_t end_b = end_rel * 4u;
        uint32_t pos = (uint32_t)((int64_t)(wb + 1) - A0) * 4u;
        uint32_t left = W;
        uint64_t *wi = p.wave_in + g0;
        uint32_t tiles_in = 1;                               // tiles waited for
        uint32_t tiles_out = 0;                              // tiles released
        mbar_wait(full_s, 0u);
        uint32_t avail = end_b < kTileBytes ? end_b : kTileBytes;        // end of the loaded region
        uint32_t rel_end = kTileBytes;                       // end of the oldest tile not yet released
        auto hop = [ring](uint32_t o) {
            uint32_t o4 = o + 4u;
            asm volatile("" : "+r"(o4));                     // (keeps the + 4 off the dependent chain)
            return o4 + 4u * *reinterpret_cast<const uint32_t *>(ring + (o & kMask));
        };
        auto record = [A0](uint32_t o) { return (uint64_t)(A0 + (int64_t)(o >> 2)); };
        while (left && pos < end_b) {
            if (pos >= rel_end) {                            // the chase has left the oldest tile: hand it back
                mbar_arrive(empty_s + 8u * (tiles_out % kLocStages));
                ++tiles_out;
                rel_end += kTileBytes;
                continue;
            }
            // four hops ahead; a hop beyond `avail` read stale data and is discarded
            const uint32_t p1 = hop(pos);
            const uint32_t p2 = hop(p1);
            const uint32_t p3 = hop(p2);
            const uint32_t p4 = hop(p3);
            if (left >= 4u && p1 > pos && p1 < avail && p2 > p1 && p2 < avail && p3 > p2 && p3 < avail && p4 > p3) {
                wi[0] = record(pos);
                wi[1] = record(p1);
                wi[2] = record(p2);
                wi[3] = record(p3);
                wi += 4;
                left -= 4u;
                pos = p4;
            } else if (pos >= avail || (avail < end_b && left >= 4u && tiles_in - tiles_out < (uint32_t)kLocStages)) {
                // out of loaded data (or too close to its end for four hops): take the next tile in.
                // (Only while a stage is free: four records that do not fit the ring together - waves of more
                // than ~10000 samples of pure noise, which only come here under DRICE_LOCATE_SCAN=0 - would have
                // the chaser wait for a tile the producer cannot issue before the chaser releases one; such
                // records are walked one hop at a time below.)
                mbar_wait(full_s + 8u * (tiles_in % kLocStages), (tiles_in / kLocStages) & 1u);
                ++tiles_in;
                const uint32_t te = tiles_in * kTileBytes;
                avail = te < end_b ? te : end_b;
            } else {
                wi[0] = record(pos);
                ++wi;
                --left;
                if (p1 <= pos) {                             // wrap: not a stream this path can hold
                    pos = end_b + 4u;
                    break;
                }
                pos = p1;
            }
        }
        if (left) s_located = W - left;
        if (left || pos != end_b) atomicOr(p.status, kErrStream);
        s_done = 1;
    } else if (threadIdx.x == 32 && !direct) {
        // ---- producer ---------------------------------------------------------------------------
        uint32_t issued = 0;
        for (uint32_t t = 0; t < ntiles; ++t) {
            const uint32_t sidx = t % kLocStages;
            if (t >= (uint32_t)kLocStages) {
                const uint32_t par = ((t / kLocStages) - 1u) & 1u;
                bool stop = false;
                while (!mbar_try_wait(empty_s + 8u * sidx, par)) {
                    if (s_done) { stop = true; break; }
                }
                if (stop) break;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            locate_issue_tile(stile_s + sidx * (kLocTileWords * 4u), full_s + 8u * sidx, p.comp,
                              A0 + (int64_t)t * kLocTileWords, p.comp_words);
            issued = t + 1;
        }
        // nothing may still be landing in shared memory when the CTA exits: wait for the last tile
        // issued into every stage
        for (uint32_t sidx = 0; sidx < (uint32_t)kLocStages; ++sidx) {
            if (issued > sidx) {
                const uint32_t last = sidx + ((issued - 1u - sidx) / kLocStages) * kLocStages;   // last tile of this stage
                mbar_wait(full_s + 8u * sidx, (last / kLocStages) & 1u);
            }
        }
    } else if (threadIdx.x >= 64) {
        // ---- the chain-independent part of the wave table -------------------------------------
        for (uint32_t w = threadIdx.x - 64; w < W; w += kLocThreads - 64) {
            const uint64_t s0 = (uint64_t)w * Lw;
            p.wave_out[g0 + w] = sb + s0;
            p.wave_n[g0 + w] = (uint32_t)((total - s0) < Lw ? (total - s0) : Lw);
        }
    }
    __syncthreads();
    // neutralise the waves that could not be located
    for (uint32_t x = s_located + threadIdx.x; x < W; x += kLocThreads) { p.wave_in[g0 + x] = wb; p.wave_out[g0 + x] = sb; p.wave_n[g0 + x] = 0; }
}

// ------------------------------------------------------------------------------------
// locate by scanning (waves of a few thousand samples: the benchmark case)
// ------------------------------------------------------------------------------------
// A record header is the word count of a wave of n samples: ceil(n (k+1) / 32) <= nwords <=
// ceil(25 n / 32).  Code words practically never fall in that range (zero runs inside the codes are
// at most k + 12 bits long, so a code word is >= 2^(19-k), far above any header of a wave shorter
// than a million samples), so the headers can be FOUND instead of chased: every SM streams part of
// the batch (scan_headers_kernel: 16 KB tiles, candidates = words in range whose record would end
// inside the chunk, kept in position order per tile), then one CTA per chunk ranks the tiles'
// candidates, fills the wave table and VERIFIES it - the table is right if and only if it starts
// at the chunk's first record, has exactly the expected number of entries and every entry + its
// word count + 1 is the next entry (the last one: the chunk's end).  If anything is off (a false
// candidate, a malformed stream) that chunk falls back to the exact serial chase through global
// memory, which also produces the error status.  The chase's one-SM-per-chunk streaming limit
// (33 GB/s per SM) is gone: the scan runs at HBM speed on all SMs.
constexpr int kScanTileWords = 4096;          // 16 KB per CTA
constexpr int kScanThreads   = 128;           // 32 words per thread
constexpr int kScanCap       = 64;            // candidates a tile can hold

struct ScanParams {
    LocateParams lp;
    uint32_t    *cnt;                         // [nchunks * max_tiles] candidates per tile (0xFFFFFFFF: overflow)
    uint32_t    *cand;                        // [nchunks * max_tiles * kScanCap] positions relative to the chunk's first word
    uint32_t     max_tiles;                   // tiles per chunk (stride of the two arrays)
    int          k;
};

struct ChunkGeom {
    uint64_t wb, we, sb, total, Lw;
    uint32_t g0, W, lo, hi;
    int64_t  A0;
};
__device__ __forceinline__ ChunkGeom chunk_geom(const LocateParams &p, uint32_t c, int k)
{
    ChunkGeom g;
    g.wb = p.chunk_word_off[c];
    g.we = p.chunk_word_off[c + 1];
    g.sb = p.chunk_sample_off[c];
    g.total = p.chunk_sample_off[c + 1] - g.sb;
    g.g0 = p.chunk_wave_off[c];
    g.W = p.chunk_wave_off[c + 1] - g.g0;
    g.Lw = p.L ? (uint64_t)p.L : g.total;
    const uint64_t n_last = g.W ? g.total - (uint64_t)(g.W - 1) * g.Lw : 0;
    const uint64_t n_min = n_last < g.Lw ? n_last : g.Lw;
    const uint64_t lo = (n_min * (uint64_t)(k + 1) + 31) / 32;
    g.lo = (uint32_t)(lo ? lo : 1);
    g.hi = (uint32_t)((25ull * g.Lw + 31) / 32);
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p.comp) >> 2) & 3u);
    g.A0 = (int64_t)((g.wb + 1 + mis) & ~3ull) - (int64_t)mis;     // 16-byte aligned address at or below the first record
    return g;
}

__global__ void __launch_bounds__(kScanThreads) scan_headers_kernel(const ScanParams sp)
{
    __shared__ uint32_t s_n;
    __shared__ uint32_t s_list[kScanCap];
    const LocateParams &p = sp.lp;
    const uint32_t c = blockIdx.y, t = blockIdx.x;
    const ChunkGeom g = chunk_geom(p, c, sp.k);
    const size_t slot = (size_t)c * sp.max_tiles + t;
    const int64_t t0 = g.A0 + (int64_t)t * kScanTileWords;
    if (g.we <= g.wb + 1 || t0 >= (int64_t)g.we) {
        if (threadIdx.x == 0) sp.cnt[slot] = 0;
        return;
    }
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint32_t span = g.hi - g.lo;
    // tiles inside the chunk (all but its first and last): no position checks, every load in flight at once
    if (t0 > (int64_t)g.wb && t0 + kScanTileWords <= (int64_t)g.we && (uint64_t)t0 + kScanTileWords <= p.comp_words) {
        constexpr int NV = kScanTileWords / 4 / kScanThreads;
        const uint4 *src = reinterpret_cast<const uint4 *>(p.comp + t0) + threadIdx.x;
        uint4 v[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) v[r] = __ldg(src + r * kScanThreads);
        const uint32_t rel0 = (uint32_t)((uint64_t)t0 - g.wb) + 4u * threadIdx.x;      // position relative to the chunk's first word
        const uint64_t room = g.we - g.wb;                                             // a record must end inside the chunk
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const uint32_t x[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
            if ((x[0] - g.lo <= span) | (x[1] - g.lo <= span) | (x[2] - g.lo <= span) | (x[3] - g.lo <= span)) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t rel = rel0 + 4u * (uint32_t)(r * kScanThreads) + (uint32_t)e;
                    if (x[e] - g.lo <= span && (uint64_t)rel + 1 + x[e] <= room) {
                        const uint32_t at = atomicAdd(&s_n, 1u);
                        if (at < (uint32_t)kScanCap) s_list[at] = rel;
                    }
                }
            }
        }
    } else
#pragma unroll
    for (int r = 0; r < kScanTileWords / 4 / kScanThreads; ++r) {
        const int64_t b = t0 + 4ll * (r * kScanThreads + (int)threadIdx.x);
        if (b >= (int64_t)g.we || b + 4 <= (int64_t)g.wb + 1) continue;
        uint32_t x[4];
        if (b >= 0 && (uint64_t)b + 4 <= p.comp_words) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.comp + b));
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (b + e >= 0 && (uint64_t)(b + e) < p.comp_words) ? p.comp[b + e] : 0xFFFFFFFFu;
        }
        // candidates are rare (one word in a few hundred): one cheap range test per word, everything else
        // (64-bit position checks) only for the words that pass it - the kernel is ALU bound otherwise
        const bool c0 = x[0] - g.lo <= span, c1 = x[1] - g.lo <= span, c2 = x[2] - g.lo <= span, c3 = x[3] - g.lo <= span;
        if (c0 | c1 | c2 | c3) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t i = b + e;
                if (x[e] - g.lo <= span && i > (int64_t)g.wb && i < (int64_t)g.we && (uint64_t)i + 1 + x[e] <= g.we) {
                    const uint32_t at = atomicAdd(&s_n, 1u);
                    if (at < (uint32_t)kScanCap) s_list[at] = (uint32_t)((uint64_t)i - g.wb);
                }
            }
        }
    }
    __syncthreads();
    const uint32_t n = s_n;
    if (n > (uint32_t)kScanCap) {
        if (threadIdx.x == 0) sp.cnt[slot] = 0xFFFFFFFFu;
        return;
    }
    if (threadIdx.x < n) {                               // position order: rank by counting (n <= 64)
        const uint32_t mine = s_list[threadIdx.x];
        uint32_t rank = 0;
        for (uint32_t m = 0; m < n; ++m) rank += s_list[m] < mine;
        sp.cand[slot * kScanCap + rank] = mine;
    }
    if (threadIdx.x == 0) sp.cnt[slot] = n;
}

constexpr int kRankThreads = 256;

__global__ void __launch_bounds__(kRankThreads) rank_headers_kernel(const ScanParams sp)
{
    __shared__ uint32_t s_warp[kRankThreads / 32];
    __shared__ uint32_t s_base, s_bad, s_located;
    const LocateParams &p = sp.lp;
    const uint32_t c = blockIdx.x;
    const ChunkGeom g = chunk_geom(p, c, sp.k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (g.we <= g.wb) {                                  // no stream at all
        if (threadIdx.x == 0) atomicOr(p.status, kErrStream);
        for (uint32_t w = threadIdx.x; w < g.W; w += kRankThreads) { p.wave_in[g.g0 + w] = g.wb; p.wave_out[g.g0 + w] = g.sb; p.wave_n[g.g0 + w] = 0; }
        return;
    }
    if (g.W == 0) {
        if (threadIdx.x == 0) {
            if (p.comp[g.wb] != (uint32_t)g.total) atomicOr(p.status, kErrTotal);
            if (g.we != g.wb + 1) atomicOr(p.status, kErrStream);
        }
        return;
    }
    if (threadIdx.x == 0) {
        s_base = 0;
        s_bad = 0;
        s_located = g.W;
        if (p.comp[g.wb] != (uint32_t)g.total) atomicOr(p.status, kErrTotal);
    }
    // the chain-independent part of the wave table
    for (uint32_t w = threadIdx.x; w < g.W; w += kRankThreads) {
        const uint64_t s0 = (uint64_t)w * g.Lw;
        p.wave_out[g.g0 + w] = g.sb + s0;
        p.wave_n[g.g0 + w] = (uint32_t)((g.total - s0) < g.Lw ? (g.total - s0) : g.Lw);
    }
    __syncthreads();
    // ---- rank: exclusive scan of the tiles' candidate counts, scatter in position order -----------
    const uint32_t ntiles = (uint32_t)(((int64_t)g.we - g.A0 + kScanTileWords - 1) / kScanTileWords);
    const size_t slot0 = (size_t)c * sp.max_tiles;
    for (uint32_t tb = 0; tb < ntiles; tb += kRankThreads) {
        const uint32_t t = tb + threadIdx.x;
        uint32_t n = t < ntiles ? sp.cnt[slot0 + t] : 0u;
        if (n == 0xFFFFFFFFu) { s_bad = 1; n = 0; }
        uint32_t inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += u;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t woff = 0, btot = 0;
#pragma unroll
        for (int w = 0; w < kRankThreads / 32; ++w) {
            if (w < warp) woff += s_warp[w];
            btot += s_warp[w];
        }
        const uint32_t first = s_base + woff + inc - n;  // wave index of the tile's first candidate
        for (uint32_t s2 = 0; s2 < n; ++s2) {
            const uint32_t w = first + s2;
            if (w < g.W) p.wave_in[g.g0 + w] = g.wb + sp.cand[(slot0 + t) * kScanCap + s2];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += btot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && s_base != g.W) s_bad = 1;
    __threadfence_block();
    __syncthreads();
    // ---- verify: the table is the chain cur += word[cur] + 1 (src/deltaRice.c:319-325) ------------
    if (!s_bad) {
        bool bad = false;
        for (uint32_t w = threadIdx.x; w < g.W; w += kRankThreads) {
            const uint64_t pos = p.wave_in[g.g0 + w];
            const uint64_t nxt = (w + 1 < g.W) ? p.wave_in[g.g0 + w + 1] : g.we;
            bad |= (pos + 1 + (uint64_t)p.comp[pos] != nxt) || (w == 0 && pos != g.wb + 1);
        }
        if (bad) s_bad = 1;
    }
    __syncthreads();
    if (s_bad) {
        // ---- fallback: the exact serial chase through global memory (also finds stream errors) ----
        if (threadIdx.x == 0) {
            uint64_t cur = g.wb + 1;
            uint32_t w = 0;
            while (w < g.W && cur < g.we) {
                p.wave_in[g.g0 + w] = cur;
                cur += (uint64_t)__ldg(p.comp + cur) + 1ull;
                ++w;
            }
            if (w < g.W) s_located = w;
            if (w < g.W || cur != g.we) atomicOr(p.status, kErrStream);
        }
        __syncthreads();
        for (uint32_t x = s_located + threadIdx.x; x < g.W; x += kRankThreads) { p.wave_in[g.g0 + x] = g.wb; p.wave_out[g.g0 + x] = g.sb; p.wave_n[g.g0 + x] = 0; }
    }
}

// ------------------------------------------------------------------------------------
// parse
// ------------------------------------------------------------------------------------
// Rice parsing is a serial chain per wave (a code's length is only known once its unary
// prefix has been read), so the parallelism is across waves: one LANE per wave, a warp takes
// 32 consecutive waves per ticket.  The kernel is bound by issue slots and by the shared-memory
// (MIO) pipe, so the inner loop is built around instructions and wavefronts per sample:
//   * one code = one lookup: the next W stream bits index a shared-memory table built for the
//     launch's k whose 4-byte entry is (delta << 16 | bits consumed).  The lane's whole decoder
//     state is ONE register S = (running sample << 16 | bit position in its ring), so a code
//     costs a single add: S += entry (the inverse delta and the bit pointer advance together;
//     the low half is folded every 16 samples so that it never carries into the sample);
//   * the table is replicated R times (entry index * R + lane % R): lanes that share a bank
//     only conflict within their group of 32 / R lanes, which cuts the wavefronts per lookup
//     from ~3.7 (random) to ~2 (R = 8, k = 2) - W and R are chosen per k to fill 32 KB;
//   * N codes per window (N * W <= 64: 6+5+5 per 16 samples for W <= 10, else 4x4): a 64-bit
//     window is rebuilt from three ring words addressed by the bit position (no window rotation,
//     no per-sample refill test), then shifted by each code's length; the N results are checked
//     for a miss ONCE (min of the entries == 0: escapes and
//     codes longer than the window), in which case the group is redone code by code with
//     count-leading-zeros;
//   * every lane of the warp advances in lock step; 16 samples pack into one 32-byte sector that
//     the lane stores straight to HBM - there is no shared-memory staging of the output;
//   * the compressed words reach the lane through a private 32-word ring in shared memory
//     ([word][lane], bank = lane: conflict free; rows 32/33 mirror rows 0/1 so that a window never
//     wraps), refilled with one 32-byte load (a full sector) that is requested a block ahead.
constexpr int      kRingWords  = 32;                 // compressed words per lane
constexpr int      kRingRows   = kRingWords + 2;     // + mirror of rows 0 and 1
constexpr uint32_t kRingBytes  = kRingRows * 128u;   // per warp
constexpr int      kChunkWords = 8;                  // one 32-byte sector per refill
constexpr uint32_t kLutBytes   = 32768;              // table incl. replication, aligned to its size
constexpr int      kLutMaxBits = 12;
constexpr uint32_t kAhead      = 16;                 // ring words guaranteed ahead at a block start:
                                                     // 16 escapes (400 bits) + the 3-word window
constexpr int kParseMaxWarps = 17;                   // per CTA; two CTAs per SM

struct Chunk8 { uint4 a, b; };
// one full 32-byte sector, L2 only: every lane streams its own record
__device__ __forceinline__ Chunk8 ldg_cg_256(const void *p)
{
    Chunk8 r;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w)
                 : "l"(p));
    return r;
}
// one full 32-byte sector per lane: partial-sector stores make L2 read the sector from DRAM first
__device__ __forceinline__ void stg_256(void *p, const uint4 &a, const uint4 &b)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
// window width / replication of the table for Rice parameter 2^k: the longest non-escape code
// has 8 + k bits; what the 32 KB do not need for the window goes into replication
__host__ __device__ __forceinline__ int lut_bits_for(int k) { return k + 8 < kLutMaxBits ? k + 8 : kLutMaxBits; }
// table entry for the W bits `idx` (MSB = next stream bit), Rice parameter 2^k:
// (delta << 16) | bits consumed for the one complete non-escape code the window starts with
// (src/deltaRice.c:161-177), 0 if there is none.
__device__ __forceinline__ uint32_t make_lut_entry(uint32_t idx, int k, int W)
{
    const uint32_t win = idx << (32 - W);
    const uint32_t q = win ? (uint32_t)__clz(win) : 32u;
    const uint32_t len = q + 1 + (uint32_t)k;
    if (q >= kEscapeQuotient || len > (uint32_t)W) return 0u;
    const uint32_t r = (win >> (32 - len)) & ((1u << k) - 1u);
    const uint32_t u = (q << k) | r;
    const int d = (u & 1u) ? -(int)((u + 1) >> 1) : (int)(u >> 1);
    return ((uint32_t)d << 16) | len;
}

// where a lane's compressed words come from (one wave): the ring holds words [.., fetched) of the
// record, counted from `ring word 0` = the 32-byte aligned block that holds the first code word
struct RingFeed {
    const uint32_t *gbase;      // 32-byte aligned address of ring word 0
    const uint32_t *comp_al;    // 32-byte aligned address at or below the stream
    uint64_t base_al, lim_al;   // ring word 0 / end of the stream, words from comp_al
    uint32_t mis;               // words between comp_al and the stream
    uint32_t safe;              // ring words [0, safe) need no bounds checks
    uint32_t ring_b;            // shared address of the lane's ring row 0
    uint32_t fetched;           // words in the ring so far (multiple of 8)

    __device__ __forceinline__ Chunk8 load_chunk(uint32_t at) const
    {
        if (at + kChunkWords <= safe) return ldg_cg_256(gbase + at);
        const uint64_t aw = base_al + at;
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = (aw + i >= mis && aw + i < lim_al) ? comp_al[aw + i] : 0u;
        Chunk8 c;
        c.a = make_uint4(e[0], e[1], e[2], e[3]);
        c.b = make_uint4(e[4], e[5], e[6], e[7]);
        return c;
    }
    __device__ __forceinline__ void store_chunk(uint32_t at, const Chunk8 &c) const
    {
        const uint32_t row = at & (kRingWords - 1);
        const uint32_t ad = ring_b + (row << 7);
        sts32(ad, c.a.x); sts32(ad + 128, c.a.y); sts32(ad + 256, c.a.z); sts32(ad + 384, c.a.w);
        sts32(ad + 512, c.b.x); sts32(ad + 640, c.b.y); sts32(ad + 768, c.b.z); sts32(ad + 896, c.b.w);
        if (row == 0) {                                      // mirror rows for windows that start in rows 30/31
            sts32(ad + kRingWords * 128, c.a.x);
            sts32(ad + kRingWords * 128 + 128, c.a.y);
        }
    }
};

// decoder state of one lane
struct LaneDec {
    uint32_t S;             // (running sample << 16) | bit position in the ring (bits 5..9 = ring row)
    uint32_t tb;            // bits folded out of S's low half (multiple of 1024)
    bool     bad;

    // shared address of the ring row that holds the current bit
    __device__ __forceinline__ uint32_t row_addr(uint32_t ring_b) const { return ((S & 0x3E0u) << 2) + ring_b; }

    // one code through count-leading-zeros (src/deltaRice.c:154-177): escapes, codes the table
    // does not hold, and the first / last few samples of a wave
    __device__ __forceinline__ void one(uint32_t ring_b, int k, uint32_t kmask)
    {
        const uint32_t a = row_addr(ring_b);
        const uint32_t win = __funnelshift_l(lds32o<128>(a), lds32o<0>(a), S);
        const uint32_t q = __clz(win);
        uint32_t u, len;
        if (q >= kEscapeQuotient) {
            bad |= (q > kEscapeQuotient);
            u = (win >> 7) & 0xFFFFu;
            len = kEscapeBits;
        } else {
            len = q + 1 + (uint32_t)k;
            u = (q << k) | ((win >> (32u - len)) & kmask);
        }
        const uint32_t h = u >> 1;
        const uint32_t d = (u & 1u) ? ~h : h;
        S += (d << 16) + len;
    }
    // keeps the low half of S below 1024 + one block, so that it never carries into the sample
    __device__ __forceinline__ void fold()
    {
        tb += S & 0xFC00u;
        S &= 0xFFFF03FFu;
    }
    __device__ __forceinline__ uint32_t bits() const { return tb + (S & 0xFFFFu); }   // from ring word 0
};

// N codes from one 64-bit window (N * W <= 64), samples FIRST .. FIRST+N-1 of a 16-sample block:
// packs them (two per register) into o[]; `carry` holds the state after an odd sample that waits
// for its partner in the next group
template <int N, int FIRST, bool IDENT>
__device__ __forceinline__ void group(LaneDec &d, uint32_t ring_b, uint32_t lutl, uint32_t imask, uint32_t ish, int k,
                                      uint32_t kmask, uint32_t (&o)[8], uint32_t &carry)
{
    const uint32_t S0 = d.S;
    const uint32_t a = d.row_addr(ring_b);
    const uint32_t w0 = lds32o<0>(a), w1 = lds32o<128>(a), w2 = lds32o<256>(a);
    uint32_t hi = __funnelshift_l(w1, w0, S0);
    uint32_t lo = __funnelshift_l(w2, w1, S0);
    uint32_t st[N];
    uint32_t S = S0, m = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint32_t e = lds32(((hi >> ish) & imask) | lutl);
        S += e;
        st[i] = S;
        m = min(m, e);
        if (i < N - 1) {
            hi = __funnelshift_l(lo, hi, e);
            if (i < N - 2) lo = __funnelshift_l(0u, lo, e);
        }
    }
    d.S = S;
    if (m == 0u) {                                           // an escape / a code longer than the window
        d.S = S0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            d.one(ring_b, k, kmask);
            st[i] = d.S;
        }
    }
    if (IDENT) {
        // no inverse delta: the sample is the decoded value itself = the step of the running state
#pragma unroll
        for (int i = N - 1; i >= 0; --i) st[i] -= i ? st[i - 1] : S0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int smp = FIRST + i;
        if (smp & 1) o[smp >> 1] = prmt(i ? st[i - 1] : carry, st[i], 0x7632);
    }
    if ((FIRST + N) & 1) carry = st[N - 1];
}

template <bool W10, bool IDENT, bool HEAVY>
__global__ void __launch_bounds__(kParseMaxWarps * 32, 2) parse_kernel(const ParseParams p)
{
    extern __shared__ __align__(16) uint32_t dsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // shared layout: the table sits at the first 32 KB boundary of the CTA's shared window (its
    // address is OR-ed, not added); the rings of the first warps fill the gap below it, the
    // others follow it
    const uint32_t dsm_s = (uint32_t)__cvta_generic_to_shared(dsm);
    const uint32_t nwarps = blockDim.x >> 5;
    const uint32_t lut_s = (dsm_s + kLutBytes - 1u) & ~(kLutBytes - 1u);
    const uint32_t nbelow = (lut_s - dsm_s) / kRingBytes;   // rings that fit below the table
    const uint32_t nabove = nwarps > nbelow ? nwarps - nbelow : 0u;
    if (lut_s + kLutBytes + nabove * kRingBytes > dsm_s + p.smem_bytes) {     // launcher and kernel disagree on the layout
        if (threadIdx.x == 0) atomicOr(p.status, kErrInternal);
        return;
    }
    const uint32_t ring_warp_s = (uint32_t)warp < nbelow ? dsm_s + (uint32_t)warp * kRingBytes
                                                         : lut_s + kLutBytes + ((uint32_t)warp - nbelow) * kRingBytes;
    const int k = p.k;
    const uint32_t kmask = (1u << k) - 1u;
    const int W = lut_bits_for(k);
    const int rsh = 13 - W;                                  // log2(replicas): 2^W entries * 4 B * R = 32 KB
    for (uint32_t i = threadIdx.x; i < kLutBytes / 4u; i += blockDim.x) sts32(lut_s + 4u * i, make_lut_entry(i >> rsh, k, W));
    __syncthreads();
    // byte offset of a table entry = (window >> ish) & imask, | the lane's replica
    const uint32_t ish = (uint32_t)(32 - W - 2 - rsh);       // = 17 for every W
    uint32_t imask = ((1u << W) - 1u) << (2 + rsh);
    asm volatile("mov.u32 %0, %0;" : "+r"(imask));          // keep it in a register (not recomputed per block)
    const uint32_t lutl = lut_s | (((uint32_t)lane & ((1u << rsh) - 1u)) << 2);

    // compressed words are fetched as 32-byte sectors: positions are words relative to the
    // 32-byte aligned address at or below p.comp
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p.comp) >> 2) & 7u);
    const uint32_t *comp_al = p.comp - mis;
    const uint64_t lim_al = p.comp_words + mis;             // end of the stream, aligned-relative
    const uint32_t ngroups = (p.nwaves + 31u) / 32u;

    // A warp's first task is fixed (warp-major over the CTAs), later ones come from the ticket counter: a batch of
    // about one task per resident warp (C2: 4794 tasks, 5032 warps) then spreads evenly over the SMs (32 or 33
    // tasks each) instead of as the atomics happen to be served (up to 34: the busiest SM sets the time).
    bool first_task = true;
    while (true) {
        uint32_t grp = 0;
        if (first_task) {
            grp = (uint32_t)warp * gridDim.x + blockIdx.x;
            first_task = false;
            if (grp >= ngroups) continue;                    // (nothing left in the fixed part: try the counter)
        } else {
            if (lane == 0) grp = nwarps * gridDim.x + atomicAdd(p.ticket, 1u);
            grp = __shfl_sync(0xffffffffu, grp, 0);
            if (grp >= ngroups) break;
        }
        const uint32_t slot = grp * 32 + lane;
        const bool active = slot < p.nwaves;
        const uint32_t g = (active && p.sort_perm) ? __ldg(p.sort_perm + slot) : slot;
        const uint32_t n = active ? __ldg(p.wave_n + g) : 0u;
        const uint64_t rec = active ? __ldg(p.wave_in + g) : 0ull;        // word index of [nwords]
        const uint64_t obase = active ? __ldg(p.wave_out + g) : 0ull;     // sample offset of the wave
        const uint32_t nwords = n ? __ldg(p.comp + rec) : 0u;
        int16_t *optr = p.out + obase;
        // A warp whose waves are heavy (more than k + 3.5 bits per sample: an escape in most groups of codes) skips the
        // table: a group with a miss is redone code by code anyway, by its lane while the others wait, so the table
        // attempt is pure overhead there.  (Warp uniform; with the density sort the warps are homogeneous.)
        const bool dens = n && (uint64_t)nwords * 256ull > (uint64_t)n * (uint32_t)(8 * k + 28);
        const bool heavy = HEAVY && __popc(__ballot_sync(0xffffffffu, dens)) >= 8;   // (HEAVY: the instance for dense batches)

        RingFeed rf;
        rf.base_al = (rec + 1 + mis) & ~7ull;               // aligned-relative index of the sector holding the first code word
        rf.gbase = comp_al + rf.base_al;
        rf.comp_al = comp_al;
        rf.lim_al = lim_al;
        rf.mis = mis;
        rf.safe = (rf.base_al >= mis)
            ? (uint32_t)(lim_al - rf.base_al > 0xFFFFFFF0ull ? 0xFFFFFFF0ull : ((lim_al - rf.base_al) & ~7ull)) : 0u;
        rf.ring_b = ring_warp_s + lane * 4u;
        rf.fetched = 0;
        const uint32_t ring_b = rf.ring_b;

        LaneDec d;
        const uint32_t start_bits = (uint32_t)((rec + 1 + mis) - rf.base_al) << 5;
        d.S = start_bits;
        d.tb = 0;
        d.bad = false;
        __syncwarp();
        if (n) {
#pragma unroll
            for (int c = 0; c < kRingWords; c += 2 * kChunkWords) {
                const Chunk8 c0 = rf.load_chunk(c), c1 = rf.load_chunk(c + kChunkWords);
                rf.store_chunk(c, c0);
                rf.store_chunk(c + kChunkWords, c1);
            }
            rf.fetched = kRingWords;
        }

        // ---- prologue: single samples up to the first 32-byte boundary of the output --------------
        // (<= 15 codes of <= 25 bits: inside the first fill)
        uint32_t left = n;
        {
            const uint32_t a = (uint32_t)((reinterpret_cast<uintptr_t>(optr) >> 1) & 15u);
            uint32_t pro = (16u - a) & 15u;
            if (pro > left) pro = left;
            left -= pro;
            for (; pro; --pro) {
                const uint32_t Sb = d.S;
                d.one(ring_b, k, kmask);
                *optr++ = (int16_t)((IDENT ? d.S - Sb : d.S) >> 16);
            }
        }
        // ---- blocks of 16 samples: 4 groups of four codes, one 32-byte store (a full sector) -------
        uint32_t nblk = left >> 4;
        left &= 15u;
        const uint32_t maxblk = __reduce_max_sync(0xffffffffu, nblk);
        Chunk8 pend;
        pend.a = pend.b = make_uint4(0, 0, 0, 0);
        bool have_pend = false;
        for (uint32_t b = 0; b <= maxblk; ++b) {
            // the last pass (b == nblk) only tops the ring up for the epilogue
            if (b <= nblk && n) {
                d.fold();
                const uint32_t cons = d.bits() >> 5;         // ring word of the current bit
                if (have_pend) {
                    rf.store_chunk(rf.fetched, pend);
                    rf.fetched += kChunkWords;
                }
                while (rf.fetched < cons + kAhead) {         // escape-heavy data outruns one sector per block
                    rf.store_chunk(rf.fetched, rf.load_chunk(rf.fetched));
                    rf.fetched += kChunkWords;
                }
                have_pend = (b < nblk) && (rf.fetched + kChunkWords <= (cons & ~(uint32_t)(kChunkWords - 1)) + kRingWords);
                if (have_pend) pend = rf.load_chunk(rf.fetched);
            }
            if (b < nblk) {
                // 16 samples as groups of N codes per window (N * W <= 64): 6+5+5 for W <= 10, else 4x4
                uint32_t o[8];
                uint32_t carry = 0;
                if (heavy) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t Sb = d.S;
                        d.one(ring_b, k, kmask);
                        const uint32_t stv = IDENT ? d.S - Sb : d.S;
                        if (i & 1) o[i >> 1] = prmt(carry, stv, 0x7632);
                        else carry = stv;
                    }
                } else if (W10) {
                    group<6, 0, IDENT>(d, ring_b, lutl, imask, ish, k, kmask, o, carry);
                    group<5, 6, IDENT>(d, ring_b, lutl, imask, ish, k, kmask, o, carry);
                    group<5, 11, IDENT>(d, ring_b, lutl, imask, ish, k, kmask, o, carry);
                } else {
                    group<4, 0, IDENT>(d, ring_b, lutl, imask, ish, k, kmask, o, carry);
                    group<4, 4, IDENT>(d, ring_b, lutl, imask, ish, k, kmask, o, carry);
                    group<4, 8, IDENT>(d, ring_b, lutl, imask, ish, k, kmask, o, carry);
                    group<4, 12, IDENT>(d, ring_b, lutl, imask, ish, k, kmask, o, carry);
                }
                stg_256(optr, make_uint4(o[0], o[1], o[2], o[3]), make_uint4(o[4], o[5], o[6], o[7]));
                optr += 16;
            }
        }
        // ---- epilogue: the last < 16 samples (the ring was topped up by the last pass) -------------
        for (; left; --left) {
            const uint32_t Sb = d.S;
            d.one(ring_b, k, kmask);
            *optr++ = (int16_t)((IDENT ? d.S - Sb : d.S) >> 16);
        }

        // the codes must end inside the last word of the record
        if (n) {
            const uint32_t used_bits = d.bits() - start_bits;
            if ((uint64_t)((used_bits + 31u) >> 5) != (uint64_t)nwords) d.bad = true;
        }
        if (d.bad) atomicOr(p.status, kErrStream);
    }
}

// ------------------------------------------------------------------------------------
// parse, small batches: one CTA per wave, parallel INSIDE the wave
// ------------------------------------------------------------------------------------
// A Rice stream is a serial chain (a code's position is the end of the previous one), so with
// one lane per wave a single HDF5 chunk of 20 waves is 20 lanes each walking 7000 codes: 0.7 ms,
// three times slower than the reference on a CPU.  For batches too small to fill the machine with
// lanes the chain is broken up instead: cut the record into runs of four words.  The bit offset at
// which the first code of a run starts (0..24: a code is at most 25 bits) determines where the
// first code of the NEXT run starts - a transition function with 25 inputs.  A warp evaluates the
// function of a run for all entry offsets at once (lane o decodes the run speculatively from offset
// o), the functions are composed along the record (a chain of one shared-memory lookup per run
// instead of one table lookup per code), and then every run is decoded independently from its now
// known entry: counts and delta sums first, a block scan for the sample indices and the inverse
// delta's running values, then the samples.  32x redundant table lookups, but spread over 512
// threads and a few SMs that would otherwise idle.
constexpr int kWideThreads  = 1024;                      // most threads a CTA uses (sizes the per-thread scratch)
constexpr int kWideRunWords = 4;                         // 128 bits per run
constexpr int kWideMaxWords = 6400;                      // ceil(25 * 8192 / 32): waves up to 8192 samples
constexpr int kWideMaxRuns  = kWideMaxWords / kWideRunWords;
constexpr uint32_t kWideMaxWaves = 896;                  // measured crossover with one lane per wave: ~900 (L = 3500) .. ~1100 (L = 7000) waves
constexpr int kWideLutBits  = 12;

struct WideCode { uint32_t len; int delta; bool valid; };
// the code at bit `pos` of the record in shared memory (words padded with four zero words)
__device__ __forceinline__ WideCode wide_decode_at(const uint32_t *words, const uint32_t *lut, uint32_t pos, int k, uint32_t kmask)
{
    const uint32_t w = pos >> 5;
    const uint32_t win = __funnelshift_l(words[w + 1], words[w], pos);
    const uint32_t e = lut[win >> (32 - kWideLutBits)];
    WideCode c;
    if (e) {
        c.len = e & 31u;
        c.delta = (int)e >> 16;
        c.valid = true;
        return c;
    }
    const uint32_t q = __clz(win);
    uint32_t u;
    c.valid = true;
    if (q == kEscapeQuotient) {
        u = (win >> 7) & 0xFFFFu;
        c.len = kEscapeBits;
    } else if (q < kEscapeQuotient) {
        c.len = q + 1 + (uint32_t)k;
        u = (q << k) | ((win >> (32u - c.len)) & kmask);
    } else {                                             // not a code (junk offset, padding, malformed stream)
        u = 0;
        c.len = 1;
        c.valid = false;
    }
    const uint32_t h = u >> 1;
    c.delta = (int)((u & 1u) ? ~h : h);
    return c;
}

// NT threads per CTA: 1024 when there are fewer waves than SMs (more warps per wave), else 512 (two CTAs per SM)
template <bool IDENT, int NT>
__global__ void __launch_bounds__(NT, NT == 1024 ? 1 : 2) parse_wide_kernel(const ParseParams p)
{
    extern __shared__ __align__(16) uint32_t wsm[];
    uint32_t *words = wsm;                                       // kWideMaxWords + 4 (zero padded: the last run may look past the record)
    uint32_t *lut = words + kWideMaxWords + 4;                   // 4096
    uint32_t *cnt = lut + (1 << kWideLutBits);                   // per thread: codes / delta sum of its runs
    uint32_t *dsum = cnt + NT;
    uint8_t *T = reinterpret_cast<uint8_t *>(dsum + NT);   // [kWideMaxRuns][32] transition functions
    uint8_t *E = T + kWideMaxRuns * 32;                          // [kWideMaxRuns + 1] entry offsets
    __shared__ uint32_t s_warp_c[NT / 32], s_warp_d[NT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = p.k;
    const uint32_t kmask = (1u << k) - 1u;
    for (uint32_t i = threadIdx.x; i < (1u << kWideLutBits); i += NT) lut[i] = make_lut_entry(i, k, kWideLutBits);

    for (uint32_t g = blockIdx.x; g < p.nwaves; g += gridDim.x) {
        __syncthreads();                                         // (previous wave done with shared memory; LUT built)
        const uint32_t n = p.wave_n[g];
        if (n == 0) continue;
        const uint64_t rec = p.wave_in[g];
        const uint32_t nw = p.comp[rec];
        int16_t *out = p.out + p.wave_out[g];
        if (nw == 0 || nw > (uint32_t)kWideMaxWords || rec + 1 + nw > p.comp_words) {
            if (threadIdx.x == 0) atomicOr(p.status, kErrStream);
            continue;
        }
        for (uint32_t i = threadIdx.x; i < nw + 4; i += NT) words[i] = i < nw ? p.comp[rec + 1 + i] : 0u;
        __syncthreads();
        const uint32_t nruns = (nw + kWideRunWords - 1) / kWideRunWords;
        // ---- transition function of every run: lane o enters at bit offset o ---------------------
        for (uint32_t r = warp; r < nruns; r += NT / 32) {
            const uint32_t end = (r + 1) * (kWideRunWords * 32);
            uint32_t pos = r * (kWideRunWords * 32) + (uint32_t)lane;
            while (pos < end) pos += wide_decode_at(words, lut, pos, k, kmask).len;
            T[r * 32 + lane] = (uint8_t)(pos - end);             // < 25
        }
        __syncthreads();
        // ---- entry offset of every run: compose along the record --------------------------------
        if (threadIdx.x == 0) {
            uint32_t e = 0;
            for (uint32_t r = 0; r < nruns; ++r) {
                E[r] = (uint8_t)e;
                e = T[r * 32 + e];
            }
        }
        __syncthreads();
        // ---- every thread takes consecutive runs: codes and delta sum first ---------------------
        const uint32_t rpt = (nruns + NT - 1) / NT;
        const uint32_t r0 = threadIdx.x * rpt, r1 = min(r0 + rpt, nruns);
        uint32_t c = 0, dsm = 0;
        for (uint32_t r = r0; r < r1; ++r) {
            const uint32_t end = (r + 1) * (kWideRunWords * 32);
            uint32_t pos = r * (kWideRunWords * 32) + E[r];
            while (pos < end) {
                const WideCode cd = wide_decode_at(words, lut, pos, k, kmask);
                pos += cd.len;
                ++c;
                dsm += (uint32_t)cd.delta;
            }
        }
        // block exclusive scan of (c, dsm)
        uint32_t ic = c, id = dsm;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t tc = __shfl_up_sync(0xffffffffu, ic, d), td = __shfl_up_sync(0xffffffffu, id, d);
            if (lane >= d) { ic += tc; id += td; }
        }
        if (lane == 31) { s_warp_c[warp] = ic; s_warp_d[warp] = id; }
        __syncthreads();
        uint32_t bc = 0, bd = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            if (w < warp) { bc += s_warp_c[w]; bd += s_warp_d[w]; }
            total += s_warp_c[w];
        }
        uint32_t idx = bc + ic - c;                              // sample index of the thread's first code
        uint32_t acc = bd + id - dsm;                            // running sample before it (src/deltaRice.c:84-89)
        bool bad = (threadIdx.x == 0 && total < n);
        // ---- the samples --------------------------------------------------------------------------
        for (uint32_t r = r0; r < r1 && idx < n; ++r) {
            const uint32_t end = (r + 1) * (kWideRunWords * 32);
            uint32_t pos = r * (kWideRunWords * 32) + E[r];
            while (pos < end && idx < n) {
                const WideCode cd = wide_decode_at(words, lut, pos, k, kmask);
                pos += cd.len;
                acc += (uint32_t)cd.delta;
                bad |= !cd.valid;
                out[idx] = (int16_t)(IDENT ? (uint32_t)cd.delta : acc);
                ++idx;
                if (idx == n) bad |= ((pos + 31u) >> 5) != nw;   // the codes must end inside the last word
            }
        }
        if (bad) atomicOr(p.status, kErrStream);
    }
}

// ------------------------------------------------------------------------------------
// parse, long waves (> 8192 samples) in small batches: several CTAs per wave
// ------------------------------------------------------------------------------------
// The reference's long-wave configurations (docs/Performance.md:27-47: 32 x 81920, 32 x 500000) and its
// DEFAULT option tuple (WaveformLength = -1: the whole chunk is one wave) give a handful of waves of up
// to millions of samples: one lane per wave would decode them on one warp.  The run decomposition of
// parse_wide_kernel carries over: the record is cut into SEGMENTS of kWideMaxWords words, one work item
// (wave, segment) per CTA, items handed out in order by a ticket.  A segment does everything that does
// not depend on its predecessor first - the transition functions of its runs, composed into the
// segment's own function (entry offset -> exit offset) for all 32 entry offsets at once - and then takes
// part in two short chains along the wave: the entry offset (one lookup in its function per link), and
// after counting its codes, the sample index and running value at its start.  A link is one global
// round trip; an item only ever waits for the item ticketed right before it, which is running or done.
constexpr int kLongGroups = 32;                          // the runs of a segment are composed in 32 groups, one per warp
constexpr unsigned long long kLongReady = 1ull << 63;

__device__ __forceinline__ unsigned long long long_wait(const unsigned long long *p)
{
    unsigned long long v;
    do {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (!(v & kLongReady)) __nanosleep(100);
    } while (!(v & kLongReady));
    return v;
}
__device__ __forceinline__ void long_post(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v | kLongReady) : "memory");
}
// words a record of n samples can have at most (25 bits per sample)
__host__ __device__ __forceinline__ uint64_t long_worst_words(uint64_t n) { return (25ull * n + 31ull) / 32ull; }

template <bool IDENT>
__global__ void __launch_bounds__(kWideThreads, 1)
parse_long_kernel(const ParseParams p, const uint32_t segs_per_wave, const uint32_t nitems, unsigned long long *const state,
                  const uint32_t seg_words)
{   // seg_words: words per segment (a multiple of kWideRunWords, <= kWideMaxWords; chosen per batch by parse_long_plan)
    constexpr int NT = kWideThreads;
    extern __shared__ __align__(16) uint32_t wsm[];
    uint32_t *words = wsm;                                       // kWideMaxWords + 4 (the last run looks past the segment)
    uint32_t *lut = words + kWideMaxWords + 4;
    uint32_t *cnt = lut + (1 << kWideLutBits);
    uint32_t *dsum = cnt + NT;
    uint8_t *T = reinterpret_cast<uint8_t *>(dsum + NT);         // [kWideMaxRuns][32]: transition functions, then the 32 paths
    uint8_t *E = T + kWideMaxRuns * 32;                          // [kWideMaxRuns + 16] entry offsets
    uint8_t *G = E + kWideMaxRuns + 16;                          // [kLongGroups][32] group functions
    uint8_t *GE = G + kLongGroups * 32;                          // [kLongGroups][32] entry of a group for every segment entry
    __shared__ uint32_t s_warp_c[NT / 32], s_warp_d[NT / 32];
    __shared__ uint32_t s_item, s_ein, s_acc0;
    __shared__ unsigned long long s_idx0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = p.k;
    const uint32_t kmask = (1u << k) - 1u;
    for (uint32_t i = threadIdx.x; i < (1u << kWideLutBits); i += NT) lut[i] = make_lut_entry(i, k, kWideLutBits);

    for (;;) {
        __syncthreads();                                         // (previous item done with shared memory; LUT built)
        if (threadIdx.x == 0) s_item = atomicAdd(p.ticket, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= nitems) break;
        const uint32_t g = item / segs_per_wave, sg = item - g * segs_per_wave;
        const uint32_t n = p.wave_n[g];
        if (n == 0) continue;
        const uint64_t rec = p.wave_in[g];
        const uint32_t nw = p.comp[rec];
        if (nw == 0 || (uint64_t)nw > long_worst_words(n) || rec + 1 + nw > p.comp_words) {   // (the same verdict in every segment)
            if (threadIdx.x == 0 && sg == 0) atomicOr(p.status, kErrStream);
            continue;
        }
        const uint64_t w0 = (uint64_t)sg * seg_words;
        if (w0 >= nw) continue;                                  // the record has fewer segments
        const uint32_t nseg = (uint32_t)((nw - w0) < (uint64_t)seg_words ? (nw - w0) : (uint64_t)seg_words);
        const bool last_seg = w0 + nseg == nw;
        int16_t *out = p.out + p.wave_out[g];
        for (uint32_t i = threadIdx.x; i < nseg + 4; i += NT) words[i] = (w0 + i < nw) ? p.comp[rec + 1 + w0 + i] : 0u;
        __syncthreads();
        const uint32_t nruns = (nseg + kWideRunWords - 1) / kWideRunWords;
        // ---- transition function of every run: lane o enters at bit offset o ---------------------
        for (uint32_t r = warp; r < nruns; r += NT / 32) {
            const uint32_t end = (r + 1) * (kWideRunWords * 32);
            uint32_t pos = r * (kWideRunWords * 32) + (uint32_t)lane;
            while (pos < end) pos += wide_decode_at(words, lut, pos, k, kmask).len;
            T[r * 32 + lane] = (uint8_t)(pos - end);             // < 25
        }
        __syncthreads();
        // ---- group functions: warp w walks its runs for all 32 entry offsets, leaving the paths in T ---
        const uint32_t gr = (nruns + kLongGroups - 1) / kLongGroups;
        {
            const uint32_t ra = warp * gr, rb = min(ra + gr, nruns);
            uint32_t e = (uint32_t)lane;
            for (uint32_t r = ra; r < rb; ++r) {
                const uint32_t x = T[r * 32 + e];
                __syncwarp();
                T[r * 32 + lane] = (uint8_t)e;                   // entry of path `lane` at run r
                __syncwarp();
                e = x;
            }
            G[warp * 32 + lane] = (uint8_t)e;
        }
        __syncthreads();
        // ---- the segment's function, then chain 1: the entry offset ------------------------------
        if (warp == 0) {
            uint32_t e = (uint32_t)lane;
            for (int w = 0; w < kLongGroups; ++w) {
                GE[w * 32 + lane] = (uint8_t)e;
                e = G[w * 32 + e];
            }
            // e = exit offset when the segment is entered at offset `lane`
            uint32_t ein = 0;
            if (lane == 0 && sg != 0) ein = (uint32_t)(long_wait(state + 2 * (size_t)(item - 1)) & 31u);
            ein = __shfl_sync(0xffffffffu, ein, 0);
            const uint32_t eout = __shfl_sync(0xffffffffu, e, (int)ein);
            if (lane == 0) {
                long_post(state + 2 * (size_t)item, eout);
                s_ein = ein;
            }
        }
        __syncthreads();
        const uint32_t ein = s_ein;
        for (uint32_t r = threadIdx.x; r < nruns; r += NT) E[r] = T[r * 32 + GE[(r / gr) * 32 + ein]];
        __syncthreads();
        // ---- every thread takes consecutive runs: codes and delta sum first ---------------------
        const uint32_t rpt = (nruns + NT - 1) / NT;
        const uint32_t r0 = threadIdx.x * rpt, r1 = min(r0 + rpt, nruns);
        uint32_t c = 0, dsm = 0;
        for (uint32_t r = r0; r < r1; ++r) {
            const uint32_t end = (r + 1) * (kWideRunWords * 32);
            uint32_t pos = r * (kWideRunWords * 32) + E[r];
            while (pos < end) {
                const WideCode cd = wide_decode_at(words, lut, pos, k, kmask);
                pos += cd.len;
                ++c;
                dsm += (uint32_t)cd.delta;
            }
        }
        uint32_t ic = c, id = dsm;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t tc = __shfl_up_sync(0xffffffffu, ic, d), td = __shfl_up_sync(0xffffffffu, id, d);
            if (lane >= d) { ic += tc; id += td; }
        }
        if (lane == 31) { s_warp_c[warp] = ic; s_warp_d[warp] = id; }
        __syncthreads();
        uint32_t bc = 0, bd = 0, total = 0, dtotal = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            if (w < warp) { bc += s_warp_c[w]; bd += s_warp_d[w]; }
            total += s_warp_c[w];
            dtotal += s_warp_d[w];
        }
        // ---- chain 2: sample index and running value at the segment's start ----------------------
        if (threadIdx.x == 0) {
            unsigned long long idx0 = 0;
            uint32_t acc0 = 0;
            if (sg != 0) {
                const unsigned long long v = long_wait(state + 2 * (size_t)(item - 1) + 1);
                idx0 = (v >> 16) & 0x7FFFFFFFFFFFull;
                acc0 = (uint32_t)(v & 0xFFFFu);
            }
            long_post(state + 2 * (size_t)item + 1, ((idx0 + total) << 16) | ((acc0 + dtotal) & 0xFFFFu));
            s_idx0 = idx0;
            s_acc0 = acc0;
        }
        __syncthreads();
        const unsigned long long idx0 = s_idx0;
        unsigned long long idx = idx0 + bc + ic - c;             // sample index of the thread's first code
        uint32_t acc = s_acc0 + bd + id - dsm;                   // running sample before it (src/deltaRice.c:84-89)
        // (a wave's last code may start in one segment and end in the next, which is then the record's last: a
        // segment that is not the last one may hold exactly the remaining codes, but not more)
        bool bad = threadIdx.x == 0 && ((last_seg && idx0 + total < n) || (!last_seg && idx0 + total > n));
        // ---- the samples --------------------------------------------------------------------------
        for (uint32_t r = r0; r < r1 && idx < n; ++r) {
            const uint32_t end = (r + 1) * (kWideRunWords * 32);
            uint32_t pos = r * (kWideRunWords * 32) + E[r];
            while (pos < end && idx < n) {
                const WideCode cd = wide_decode_at(words, lut, pos, k, kmask);
                pos += cd.len;
                acc += (uint32_t)cd.delta;
                bad |= !cd.valid;
                out[idx] = (int16_t)(IDENT ? (uint32_t)cd.delta : acc);
                ++idx;
                if (idx == n) bad |= (uint64_t)((pos + 31u) >> 5) != (uint64_t)nw - w0;   // the codes must end inside the record's last word
            }
        }
        if (bad) atomicOr(p.status, kErrStream);
    }
}

// Segments of parse_long_kernel for a batch.  The host only knows the WORST record size (25 bits per sample; real
// records are about a quarter of it), so the segment is sized for ~6 worst-case items per SM - one to two real
// ones: long waves get segments of up to kWideMaxWords, a handful of SHORT waves (one HDF5 chunk of the README's
// 20 x 7000) is cut into segments of >= 256 words as well instead of one CTA per wave (parse_wide_kernel).
struct LongPlan { uint32_t seg_words, segs; };
LongPlan parse_long_plan(uint32_t nwaves, uint32_t max_n, int sms)
{
    const uint64_t worst = long_worst_words(max_n);
    uint64_t sw = (worst * (uint64_t)nwaves + 6ull * (uint64_t)sms - 1) / (6ull * (uint64_t)sms);
    sw = (sw + 63ull) & ~63ull;
    if (sw < 256ull) sw = 256ull;
    if (sw > (uint64_t)kWideMaxWords) sw = (uint64_t)kWideMaxWords;
    LongPlan pl;
    pl.seg_words = (uint32_t)sw;
    pl.segs = (uint32_t)((worst + sw - 1) / sw);
    return pl;
}
// does the batch take parse_long_kernel?  (waves longer than 8192 samples always, shorter ones when the plan
// cuts them into at least 16 worst-case segments, i.e. about four real ones)
bool parse_long_applies(uint32_t nwaves, uint32_t max_n, int sms, LongPlan *pl)
{
    static long wide = -1;                                   // DRICE_PARSE_WIDE=<max waves> (0: always one lane per wave)
    if (wide < 0) {
        const char *e = getenv("DRICE_PARSE_WIDE");
        wide = e ? atol(e) : (long)kWideMaxWaves;
    }
    static const int short_env = [] { const char *v = getenv("DRICE_PARSE_LONG_SHORT"); return v ? atoi(v) : 1; }();
    if ((long)nwaves > wide || nwaves == 0) return false;
    *pl = parse_long_plan(nwaves, max_n, sms);
    if (max_n > 8192u) return true;
    return short_env != 0 && pl->segs >= 16u;                // (20 x 7000: 74 -> 52 us; from ~50 waves on one CTA per wave is as fast)
}

// warps per SM: as many as fit, trimmed so that the last round of warp tasks is nearly full
int pick_parse_warps(uint32_t ngroups, int sms, int max_warps)
{
    int best = max_warps;
    double best_eff = 0.0;
    for (int w = max_warps; w >= max_warps / 2 && w >= 4; --w) {
        const double slots = (double)w * sms;
        const double rounds = (double)ngroups / slots;
        const double eff = rounds / (double)((uint64_t)((ngroups + (uint64_t)slots - 1) / (uint64_t)slots));
        if (eff > best_eff + 0.02) { best_eff = eff; best = w; }
    }
    return best;
}

// ------------------------------------------------------------------------------------
// waves sorted by density for the lane parser
// ------------------------------------------------------------------------------------
// The 32 lanes of a warp advance in lock step, and a group of codes with an escape (or any code longer than
// the table's window) is redone by its lane while the others wait: a warp is as slow as its heaviest wave.
// When a batch mixes noise levels (BASELINE config C4), waves of similar bits per sample are put into the
// same warps instead: a counting sort into 8 buckets of (record words * 32 / samples) relative to k - a
// histogram pass and a scatter pass over the wave table (order inside a bucket is arbitrary, tiles of
// consecutive waves stay together).  Only the ORDER in which waves are decoded changes.
constexpr int kSortBuckets = 8;
constexpr int kSortThreads = 256, kSortPerThread = 4;
__device__ __forceinline__ uint32_t density_bucket(const ParseParams &p, uint32_t g)
{
    const uint32_t n = __ldg(p.wave_n + g);
    const uint64_t rec = __ldg(p.wave_in + g);
    if (n == 0 || rec >= p.comp_words) return 0;
    const uint64_t nw = __ldg(p.comp + rec);
    const uint64_t x8 = nw * 256ull / n;                     // bits per sample, in eighths
    const uint32_t k8 = (uint32_t)p.k * 8u;
    const uint32_t t[kSortBuckets - 1] = {k8 + 20u, k8 + 28u, k8 + 40u, k8 + 56u, k8 + 80u, k8 + 112u, k8 + 152u};
    uint32_t b = 0;
#pragma unroll
    for (int i = 0; i < kSortBuckets - 1; ++i) b += x8 > t[i];
    return b;
}
__global__ void __launch_bounds__(kSortThreads) wave_hist_kernel(const ParseParams p)
{
    __shared__ uint32_t h[kSortBuckets];
    if (threadIdx.x < kSortBuckets) h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t g0 = blockIdx.x * (kSortThreads * kSortPerThread);
#pragma unroll
    for (int i = 0; i < kSortPerThread; ++i) {
        const uint32_t g = g0 + i * kSortThreads + threadIdx.x;
        if (g < p.nwaves) atomicAdd(&h[density_bucket(p, g)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kSortBuckets && h[threadIdx.x]) atomicAdd(p.sort_counters + threadIdx.x, h[threadIdx.x]);
}
__global__ void __launch_bounds__(kSortThreads) wave_scatter_kernel(const ParseParams p)
{
    __shared__ uint32_t h[kSortBuckets], base[kSortBuckets];
    if (threadIdx.x < kSortBuckets) h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t g0 = blockIdx.x * (kSortThreads * kSortPerThread);
    uint32_t b[kSortPerThread], r[kSortPerThread];
#pragma unroll
    for (int i = 0; i < kSortPerThread; ++i) {
        const uint32_t g = g0 + i * kSortThreads + threadIdx.x;
        b[i] = 0; r[i] = 0;
        if (g < p.nwaves) { b[i] = density_bucket(p, g); r[i] = atomicAdd(&h[b[i]], 1u); }
    }
    __syncthreads();
    if (threadIdx.x < kSortBuckets) {
        uint32_t start = 0;                                  // first slot of the bucket: buckets before it (histogram pass)
        for (int j = 0; j < (int)threadIdx.x; ++j) start += p.sort_counters[j];
        base[threadIdx.x] = start + (h[threadIdx.x] ? atomicAdd(p.sort_counters + kSortBuckets + threadIdx.x, h[threadIdx.x]) : 0u);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortPerThread; ++i) {
        const uint32_t g = g0 + i * kSortThreads + threadIdx.x;
        if (g < p.nwaves) p.sort_perm[base[b[i]] + r[i]] = g;
    }
}

int launch_parse_impl(const ParseParams &p, cudaStream_t st)
{
    const int g_dec_sms = device_sm_count();
    static DeviceOnce attr_set;                          // (function attributes belong to a device)
    if (attr_set.first()) {
        auto opt_in = [](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        };
        opt_in(parse_kernel<true, false, false>); opt_in(parse_kernel<false, false, false>);
        opt_in(parse_kernel<true, true, false>);  opt_in(parse_kernel<false, true, false>);
        opt_in(parse_kernel<true, false, true>);  opt_in(parse_kernel<false, false, true>);
        opt_in(parse_kernel<true, true, true>);   opt_in(parse_kernel<false, true, true>);
    }
    // small batches: one CTA per wave (parallel inside the wave) instead of one lane per wave
    {
        static long wide = -1;                               // DRICE_PARSE_WIDE=<max waves> (0: always one lane per wave)
        if (wide < 0) {
            const char *e = getenv("DRICE_PARSE_WIDE");
            wide = e ? atol(e) : (long)kWideMaxWaves;
        }
        LongPlan pl;
        if (p.long_state && parse_long_applies(p.nwaves, p.max_n, g_dec_sms, &pl)) {
            static DeviceOnce lattr;
            const size_t lsmem = (size_t)(kWideMaxWords + 4 + (1 << kWideLutBits) + 2 * kWideThreads) * 4 +
                                 (size_t)kWideMaxRuns * 32 + kWideMaxRuns + 16 + 2 * kLongGroups * 32 + 16;
            if (lattr.first()) {
                cudaFuncSetAttribute(parse_long_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem);
                cudaFuncSetAttribute(parse_long_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem);
            }
            const uint64_t nitems = (uint64_t)p.nwaves * pl.segs;
            const uint32_t lgrid = nitems < (uint64_t)g_dec_sms ? (uint32_t)nitems : (uint32_t)g_dec_sms;
            if (p.identity) parse_long_kernel<true><<<lgrid, kWideThreads, lsmem, st>>>(p, pl.segs, (uint32_t)nitems, p.long_state, pl.seg_words);
            else            parse_long_kernel<false><<<lgrid, kWideThreads, lsmem, st>>>(p, pl.segs, (uint32_t)nitems, p.long_state, pl.seg_words);
            return 1;
        }
        if ((long)p.nwaves <= wide && p.max_n <= 8192u) {
            static DeviceOnce wattr;
            const size_t wsmem = (size_t)(kWideMaxWords + 4 + (1 << kWideLutBits) + 2 * kWideThreads) * 4 +
                                 (size_t)kWideMaxRuns * 32 + kWideMaxRuns + 16;
            if (wattr.first()) {
                cudaFuncSetAttribute(parse_wide_kernel<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
                cudaFuncSetAttribute(parse_wide_kernel<true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
                cudaFuncSetAttribute(parse_wide_kernel<false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
                cudaFuncSetAttribute(parse_wide_kernel<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
            }
            if (p.nwaves <= (uint32_t)g_dec_sms) {           // fewer waves than SMs: the biggest CTA per wave
                if (p.identity) parse_wide_kernel<true, 1024><<<p.nwaves, 1024, wsmem, st>>>(p);
                else            parse_wide_kernel<false, 1024><<<p.nwaves, 1024, wsmem, st>>>(p);
            } else {
                const uint32_t wgrid = p.nwaves < (uint32_t)(2 * g_dec_sms) ? p.nwaves : (uint32_t)(2 * g_dec_sms);
                if (p.identity) parse_wide_kernel<true, 512><<<wgrid, 512, wsmem, st>>>(p);
                else            parse_wide_kernel<false, 512><<<wgrid, 512, wsmem, st>>>(p);
            }
            return 1;
        }
    }
    const uint32_t ngroups = (p.nwaves + 31u) / 32u;
    static int max_warps = 0;
    if (!max_warps) {
        const char *e = getenv("DRICE_DEC_WARPS");
        max_warps = e ? atoi(e) : kParseMaxWarps;
        if (max_warps < 1) max_warps = 1;
        if (max_warps > kParseMaxWarps) max_warps = kParseMaxWarps;
    }
    // two CTAs per SM; warps per CTA trimmed so that the last round of warp tasks is nearly full
    const int ctas = 2 * g_dec_sms;
    int warps = pick_parse_warps(ngroups, ctas, max_warps);
    uint32_t grid = (uint32_t)ctas;
    if ((uint64_t)grid * warps > ngroups) {
        // small batch: spread the groups over the CTAs
        grid = ngroups < (uint32_t)ctas ? ngroups : (uint32_t)ctas;
        warps = (int)((ngroups + grid - 1) / grid);
        if (warps < 1) warps = 1;
    }
    // shared layout: [pad to 32 KB] table | rings.  The dynamic area starts right after the 1 KB the
    // system reserves per CTA (the kernel checks the assumption).
    size_t off = 1024;
    const size_t lut_at = (off + kLutBytes - 1) & ~(size_t)(kLutBytes - 1);
    const size_t nbelow = (lut_at - off) / kRingBytes;
    const size_t nabove = (size_t)warps > nbelow ? (size_t)warps - nbelow : 0;
    off = lut_at + kLutBytes + nabove * kRingBytes;
    ParseParams pp = p;
    pp.smem_bytes = (uint32_t)(off - 1024);
    int launches = 1;
    if (pp.sort_perm && pp.sort_counters) {
        const uint32_t sgrid = (p.nwaves + kSortThreads * kSortPerThread - 1) / (kSortThreads * kSortPerThread);
        wave_hist_kernel<<<sgrid, kSortThreads, 0, st>>>(pp);
        wave_scatter_kernel<<<sgrid, kSortThreads, 0, st>>>(pp);
        launches += 2;
    } else {
        pp.sort_perm = nullptr;
    }
    if (getenv("DRICE_DEBUG")) {
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, parse_kernel<true, false, false>, warps * 32, pp.smem_bytes);
        fprintf(stderr, "parse: grid %u warps %d smem %u occ %d\n", grid, warps, pp.smem_bytes, occ);
    }
    const bool w10 = lut_bits_for(p.k) <= 10;
    // dense batches (the ones that are sorted): the instance whose heavy warps skip the table
    static const int heavy_env = [] { const char *v = getenv("DRICE_DEC_HEAVY"); return v ? atoi(v) : 1; }();
    const bool hv = pp.heavy_ok != 0 && heavy_env != 0;
    auto go = [&](auto kern) { kern<<<grid, warps * 32, pp.smem_bytes, st>>>(pp); };
    if (hv) {
        if (p.identity) { if (w10) go(parse_kernel<true, true, true>); else go(parse_kernel<false, true, true>); }
        else            { if (w10) go(parse_kernel<true, false, true>); else go(parse_kernel<false, false, true>); }
    } else {
        if (p.identity) { if (w10) go(parse_kernel<true, true, false>); else go(parse_kernel<false, true, false>); }
        else            { if (w10) go(parse_kernel<true, false, false>); else go(parse_kernel<false, false, false>); }
    }
    return launches;
}

}  // namespace

// zeroed state (bytes) parse_long_kernel needs for a batch; 0 when the batch does not take that path
size_t parse_long_state_bytes(uint32_t nwaves, uint32_t max_n)
{
    LongPlan pl;
    if (!parse_long_applies(nwaves, max_n, device_sm_count(), &pl)) return 0;
    const uint64_t items = (uint64_t)nwaves * pl.segs;
    if (items >= (1ull << 31)) return 0;
    return (size_t)items * 16;
}

int launch_locate(const LocateParams &p, cudaStream_t st)
{
    if (p.nchunks == 0) return 0;
    static DeviceOnce attr_set;
    const size_t smem = (size_t)(kLocStages * kLocTileWords) * sizeof(uint32_t);
    if (attr_set.first()) cudaFuncSetAttribute(locate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    locate_kernel<<<p.nchunks, kLocThreads, smem, st>>>(p);
    return 1;
}

// Direct chase: one warp per chunk, lane 0 walks cur += word[cur] + 1 straight through global memory (one
// dependent 32-byte read per record, ~0.8 us a hop), the other lanes fill the chain-independent part of the wave
// table.  All chunks chase at once, so the time is (waves per chunk) hops whatever the batch size, while the
// scan reads the whole stream: for very large batches (C4: 32 GiB of stream, 2321 chunks of 2000 waves) the
// chase is the cheaper one (~2 ms against 4.9 ms).
__global__ void __launch_bounds__(32) chase_direct_kernel(const LocateParams p, const int k)
{
    const uint32_t c = blockIdx.x;
    const int lane = threadIdx.x;
    const ChunkGeom g = chunk_geom(p, c, k);
    if (g.we <= g.wb) {                                  // no stream at all
        if (lane == 0) atomicOr(p.status, kErrStream);
        for (uint32_t w = lane; w < g.W; w += 32) { p.wave_in[g.g0 + w] = g.wb; p.wave_out[g.g0 + w] = g.sb; p.wave_n[g.g0 + w] = 0; }
        return;
    }
    if (g.W == 0) {
        if (lane == 0) {
            if (p.comp[g.wb] != (uint32_t)g.total) atomicOr(p.status, kErrTotal);
            if (g.we != g.wb + 1) atomicOr(p.status, kErrStream);
        }
        return;
    }
    uint32_t located = g.W;
    if (lane == 0) {
        if (p.comp[g.wb] != (uint32_t)g.total) atomicOr(p.status, kErrTotal);
        uint64_t cur = g.wb + 1;
        uint32_t w = 0;
        while (w < g.W && cur < g.we) {                  // (src/deltaRice.c:319-325)
            p.wave_in[g.g0 + w] = cur;
            cur += (uint64_t)__ldg(p.comp + cur) + 1ull;
            ++w;
        }
        if (w < g.W) located = w;
        if (w < g.W || cur != g.we) atomicOr(p.status, kErrStream);
    }
    located = __shfl_sync(0xffffffffu, located, 0);
    for (uint32_t w = lane; w < located; w += 32) {
        const uint64_t s0 = (uint64_t)w * g.Lw;
        p.wave_out[g.g0 + w] = g.sb + s0;
        p.wave_n[g.g0 + w] = (uint32_t)((g.total - s0) < g.Lw ? (g.total - s0) : g.Lw);
    }
    for (uint32_t x = located + lane; x < g.W; x += 32) { p.wave_in[g.g0 + x] = g.wb; p.wave_out[g.g0 + x] = g.sb; p.wave_n[g.g0 + x] = 0; }
}
// the direct chase instead of scan + rank: hops of the longest chunk (~0.8 us each) against reading the stream
// (~5 TB/s), with a factor of two in favour of the scan
bool locate_direct_applies(uint64_t comp_words, uint64_t max_chunk_waves, size_t nchunks)
{
    static const int mode = [] { const char *e = getenv("DRICE_LOCATE_DIRECT"); return e ? atoi(e) : -1; }();   // 1 / 0: always / never
    if (mode >= 0) return mode != 0;
    // measured: 0.45 us a hop (2000 hops: 0.9 ms) against ~5.3 TB/s for the scan
    return nchunks >= 512 && comp_words > max_chunk_waves * 1500000ull;
}
int launch_locate_direct(const LocateParams &p, int k, cudaStream_t st)
{
    if (p.nchunks == 0) return 0;
    chase_direct_kernel<<<p.nchunks, 32, 0, st>>>(p, k);
    return 1;
}

// scan + rank instead of the chase: when the waves are long enough for a tile to hold all its
// headers (and short enough for the scan to make sense).  `max_chunk_words`: longest chunk stream.
bool locate_scan_applies(uint32_t L, int k, uint64_t max_chunk_words, size_t nchunks, size_t *scratch_bytes)
{
    static int mode = -1;                                // DRICE_LOCATE_SCAN=0 forces the chase
    if (mode < 0) {
        const char *e = getenv("DRICE_LOCATE_SCAN");
        mode = e ? atoi(e) : 1;
    }
    if (!mode || L == 0 || (uint64_t)L >= kLocDirectL || k < 0) return false;
    const uint64_t lo_full = ((uint64_t)L * (uint64_t)(k + 1) + 31) / 32;
    if ((lo_full + 1) * (uint64_t)(kScanCap - 2) < (uint64_t)kScanTileWords) return false;
    const uint64_t max_tiles = (max_chunk_words + 3) / kScanTileWords + 2;
    const uint64_t bytes = (uint64_t)nchunks * max_tiles * (kScanCap + 1) * 4;
    if (max_tiles > 65535 || nchunks > 0x7fffffffull || bytes > (512ull << 20)) return false;
    *scratch_bytes = (size_t)bytes;
    return true;
}

int launch_locate_scan(const LocateParams &p, int k, uint64_t max_chunk_words, uint64_t max_chunk_waves, void *scratch, cudaStream_t st)
{
    if (p.nchunks == 0) return 0;
    (void)max_chunk_waves;
    ScanParams sp;
    sp.lp = p;
    sp.k = k;
    sp.max_tiles = (uint32_t)((max_chunk_words + 3) / kScanTileWords + 2);
    sp.cnt = (uint32_t *)scratch;
    sp.cand = sp.cnt + (size_t)p.nchunks * sp.max_tiles;
    // (y = chunk: grids of more than 65535 chunks go in slices)
    for (uint32_t c0 = 0; c0 < p.nchunks; c0 += 65535u) {
        ScanParams q = sp;
        const uint32_t nc = p.nchunks - c0 < 65535u ? p.nchunks - c0 : 65535u;
        q.lp.chunk_word_off += c0;
        q.lp.chunk_sample_off += c0;
        q.lp.chunk_wave_off += c0;
        q.lp.nchunks = nc;
        q.cnt += (size_t)c0 * sp.max_tiles;
        q.cand += (size_t)c0 * sp.max_tiles * kScanCap;
        scan_headers_kernel<<<dim3(sp.max_tiles, nc), kScanThreads, 0, st>>>(q);
    }
    rank_headers_kernel<<<p.nchunks, kRankThreads, 0, st>>>(sp);
    return 2;
}

int launch_parse(const ParseParams &p, int store_bytes, cudaStream_t st)
{
    (void)store_bytes;          // every alignment takes the same path: 16-byte aligned blocks per lane
    if (p.nwaves == 0) return 0;
    if (p.k < 0 || p.k > 15) return -1;
    return launch_parse_impl(p, st);
}

}  // namespace drice
