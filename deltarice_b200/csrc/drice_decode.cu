// drice_decode.cu — Rice DECODE + inverse delta for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   readWholeCompressedByteString   src/deltaRice.c:301-341 (header walk :319-325)
//   perWaveDecompression            src/deltaRice.c:293-297
//   decompressWithRiceCoding        src/deltaRice.c:138-189
//   decodeWaveform delta branch     src/deltaRice.c:78-90
//
// Two kernels:
//   locate_kernel  one CTA per chunk.  The stream has no index, only the chain
//                  cur += word[cur] + 1 (:319-325); chasing it through HBM would cost one
//                  DRAM round trip per wave, so the CTA streams the chunk through shared
//                  memory (cp.async, double buffered) and one thread chases the chain at
//                  shared-memory latency, writing the per-wave table (record position,
//                  output position, sample count).  Tiles that hold no header are skipped.
//   parse_kernel   one THREAD per wave, 32 waves per warp: Rice parsing is a serial chain
//                  per wave, so the parallelism is across waves.  Each lane streams its
//                  record through a private shared-memory ring (128-bit loads issued one
//                  group ahead), finds the unary terminator with one count-leading-zeros on
//                  a funnel-shifted 32-bit window, rebuilds the sample with a running sum
//                  (inverse delta, wraps mod 2^16) and writes it to a per-warp shared tile
//                  that the warp then stores to HBM row by row, coalesced.
#include "drice_kernels.cuh"

#include <cstdio>
#include <cstdlib>

namespace drice {

namespace {

// ------------------------------------------------------------------------------------
// locate
// ------------------------------------------------------------------------------------
// The stream has no index, only the chain cur += word[cur] + 1 (src/deltaRice.c:319-325), and a
// hop through HBM costs a DRAM round trip.  One CTA per chunk therefore streams the chunk through
// shared memory (cp.async, three tiles in flight) while ONE thread chases the chain at shared-
// memory latency, writing the record positions of the tile to a list; the whole CTA then turns
// the list into wave table entries.  Records longer than the pipeline are jumped over.
constexpr int kLocThreads   = 256;
constexpr int kLocTileWords = 8192;          // 32 KB per stage
constexpr int kLocStages    = 3;
constexpr int kLocListMax   = kLocTileWords;       // a record is at least its [nwords] word

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// loads words [A, A+kLocTileWords) into `dst` ((comp + A) is 16-byte aligned; A may be negative
// by up to 3 words when comp itself is not 16-byte aligned); words outside [0, limit) are skipped.
__device__ __forceinline__ void locate_load_tile(uint32_t *dst, const uint32_t *comp, int64_t A, uint64_t limit)
{
    for (int v = threadIdx.x; v < kLocTileWords / 4; v += kLocThreads) {
        const int64_t w = A + 4ll * v;
        if (w >= 0 && (uint64_t)w + 4 <= limit) {
            cp_async16(dst + 4 * v, comp + w);
        } else {
            for (int e = 0; e < 4; ++e)
                if (w + e >= 0 && (uint64_t)(w + e) < limit) dst[4 * v + e] = comp[w + e];
        }
    }
}

__global__ void __launch_bounds__(kLocThreads) locate_kernel(const LocateParams p)
{
    extern __shared__ __align__(16) uint32_t stile[];   // kLocStages * kLocTileWords | list of kLocListMax
    uint32_t *s_list = stile + kLocStages * kLocTileWords;      // tile-relative record positions
    __shared__ uint64_t s_cur;
    __shared__ uint32_t s_wave, s_cnt;
    __shared__ int s_state;                              // 0 = next tile, 1 = done, 2 = jump to s_cur
    const uint32_t c = blockIdx.x;
    const uint64_t wb = p.chunk_word_off[c], we = p.chunk_word_off[c + 1];
    const uint64_t sb = p.chunk_sample_off[c], se = p.chunk_sample_off[c + 1];
    const uint32_t g0 = p.chunk_wave_off[c];
    const uint32_t W = p.chunk_wave_off[c + 1] - g0;    // waves expected from the caller's sizes
    const uint64_t total = se - sb;
    const uint64_t Lw = p.L ? (uint64_t)p.L : total;

    if (we <= wb) {                                      // no stream at all
        if (threadIdx.x == 0) atomicOr(p.status, kErrStream);
        return;
    }
    if (threadIdx.x == 0) {
        if (p.comp[wb] != (uint32_t)total) atomicOr(p.status, kErrTotal);
        if (W == 0 && we != wb + 1) atomicOr(p.status, kErrStream);
        s_cur = wb + 1;
        s_wave = 0;
        s_state = (W == 0) ? 1 : 0;
    }
    __syncthreads();
    if (s_state == 1) return;

    // misalignment of the global address: tiles start at word indices A with (comp + A) 16-byte aligned
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p.comp) >> 2) & 3u);
    auto align_down = [mis](uint64_t w) { return (int64_t)((w + mis) & ~3ull) - (int64_t)mis; };

    int64_t A = align_down(wb + 1);                      // first word of the oldest tile in flight
    int head = 0;                                        // its stage
    auto refill_all = [&]() {
#pragma unroll
        for (int sidx = 0; sidx < kLocStages; ++sidx) {
            const int64_t At = A + (int64_t)sidx * kLocTileWords;
            if ((uint64_t)(At < 0 ? 0 : At) < we) locate_load_tile(stile + ((head + sidx) % kLocStages) * kLocTileWords, p.comp, At, we);
            cp_async_commit();
        }
    };
    refill_all();
    while (true) {
        cp_async_wait<kLocStages - 1>();                 // the oldest tile has landed
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t *t = stile + head * kLocTileWords;
            uint64_t cur = s_cur;
            uint32_t w = s_wave, cnt = 0;
            const uint64_t tile_end = (uint64_t)(A + kLocTileWords);
            const uint64_t stop = tile_end < we ? tile_end : we;
            if (cur < stop) {
                // tile-relative 32-bit chase: the dependent chain is one LDS + one add per hop
                uint32_t rel = (uint32_t)((int64_t)cur - A);
                const uint32_t stop_rel = (uint32_t)((int64_t)stop - A);
                const uint32_t wleft = W - w;
                while (cnt < wleft && rel < stop_rel) {
                    s_list[cnt++] = rel;
                    rel += t[rel] + 1u;
                }
                w += cnt;
                cur = (uint64_t)(A + (int64_t)rel);
            }
            s_cur = cur;
            s_cnt = cnt;
            int state = 0;
            if (w == W) {
                if (cur != we) atomicOr(p.status, kErrStream);
                state = 1;
            } else if (cur >= we) {
                atomicOr(p.status, kErrStream);
                // neutralise the waves that could not be located
                for (uint32_t x = w; x < W; ++x) { p.wave_in[g0 + x] = wb; p.wave_out[g0 + x] = sb; p.wave_n[g0 + x] = 0; }
                state = 1;
            } else if (cur >= (uint64_t)(A + (int64_t)kLocStages * kLocTileWords)) {
                state = 2;                               // a record longer than everything in flight
            }
            s_state = state;
        }
        __syncthreads();
        // the tile's records -> wave table, by everybody
        {
            const uint32_t cnt = s_cnt, w0 = s_wave;
            for (uint32_t i = threadIdx.x; i < cnt; i += kLocThreads) {
                const uint32_t w = w0 + i;
                const uint64_t s0 = (uint64_t)w * Lw;
                p.wave_in[g0 + w] = (uint64_t)(A + (int64_t)s_list[i]);
                p.wave_out[g0 + w] = sb + s0;
                p.wave_n[g0 + w] = (uint32_t)((total - s0) < Lw ? (total - s0) : Lw);
            }
        }
        const int state = s_state;
        __syncthreads();                                 // list and tile are free again
        if (threadIdx.x == 0) s_wave += s_cnt;
        if (state == 1) break;
        if (state == 2) {                                // restart the pipeline at the record's header
            cp_async_wait<0>();
            __syncthreads();
            A = align_down(s_cur);
            head = 0;
            refill_all();
        } else {                                         // reuse the stage for the tile after the ones in flight
            const int64_t An = A + (int64_t)kLocStages * kLocTileWords;
            if ((uint64_t)An < we) locate_load_tile(stile + head * kLocTileWords, p.comp, An, we);
            cp_async_commit();
            A += kLocTileWords;
            head = (head + 1) % kLocStages;
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------
// parse
// ------------------------------------------------------------------------------------
// Rice parsing is a serial chain per wave (a code's length is only known once its unary
// prefix has been read), so the parallelism is across waves: one LANE per wave, a warp takes
// 32 consecutive waves per ticket.  The kernel is bound by issue slots and by the shared-memory
// (MIO) pipe, so the inner loop is built to touch shared memory as little as possible:
//   * one code = one lookup: the next 12 stream bits index a shared-memory table built for the
//     launch's k whose 4-byte entry is (delta << 16 | bits consumed); escapes and codes longer
//     than the window miss (entry 0) and take a count-leading-zeros path;
//   * a step decodes exactly TWO samples (two chained lookups in a 96-bit register window), so
//     every lane of the warp advances in lock step and the two 16-bit results pack into one
//     register; four steps fill a 16-byte block that the lane stores straight to HBM - there is
//     no shared-memory staging of the output at all;
//   * the compressed words reach the lane through a private 16-word ring in shared memory
//     ([word][lane], bank = lane: conflict free), refilled with one 16-byte load per block that is
//     requested a block ahead.
constexpr int kLutBits     = 12;
constexpr int kLutSize     = 1 << kLutBits;
constexpr int kRingWords   = 32;                 // compressed words per lane (4 KB per warp)
constexpr int kParseMaxWarps = 17;               // per CTA; two CTAs per SM

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
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ldg_cg_u4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"      // L2 only: every lane streams its own record
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// one full 32-byte sector per lane: partial-sector stores make L2 read the sector from DRAM first
__device__ __forceinline__ void stg_256(void *p, const uint4 &a, const uint4 &b)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
// table entry for the kLutBits bits `idx` (MSB = next stream bit), Rice parameter 2^k:
// (delta << 16) | bits consumed for the one complete non-escape code the window starts with
// (src/deltaRice.c:161-177), 0 if there is none.
__device__ __forceinline__ uint32_t make_lut_entry(uint32_t idx, int k)
{
    const uint32_t win = idx << (32 - kLutBits);
    const uint32_t q = win ? (uint32_t)__clz(win) : 32u;
    const uint32_t len = q + 1 + (uint32_t)k;
    if (q >= kEscapeQuotient || len > (uint32_t)kLutBits) return 0u;
    const uint32_t r = (win >> (32 - len)) & ((1u << k) - 1u);
    const uint32_t u = (q << k) | r;
    const int d = (u & 1u) ? -(int)((u + 1) >> 1) : (int)(u >> 1);
    return ((uint32_t)d << 16) | len;
}

// where a lane's compressed words come from (one wave)
struct RingFeed {
    const uint32_t *gbase;      // 16-byte aligned address of chunk 0
    const uint32_t *comp_al;    // aligned address at or below the stream
    uint64_t base_al, lim_al;   // chunk 0 / end of the stream, words from comp_al
    uint32_t mis;               // words between comp_al and the stream
    uint32_t safe;              // chunks [0, safe) need no bounds checks
    uint32_t ring_b;            // shared address of the lane's ring word 0 (region aligned to its size)
    uint32_t fetched;           // words in the ring so far (multiple of 4), from chunk 0

    __device__ __forceinline__ uint4 load_chunk(uint32_t at) const
    {
        if (at + 4 <= safe) return ldg_cg_u4(gbase + at);
        const uint64_t aw = base_al + at;
        uint32_t e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = (aw + i >= mis && aw + i < lim_al) ? comp_al[aw + i] : 0u;
        return make_uint4(e[0], e[1], e[2], e[3]);
    }
    __device__ __forceinline__ void store_chunk(uint32_t at, const uint4 &c) const
    {
        const uint32_t ad = ((at << 7) & ((kRingWords << 7) - 128u)) | ring_b;
        sts32(ad, c.x); sts32(ad + 128, c.y); sts32(ad + 256, c.z); sts32(ad + 384, c.w);
    }
    __device__ __forceinline__ uint32_t word(uint32_t x) const
    {
        return lds32(((x << 7) & ((kRingWords << 7) - 128u)) | ring_b);
    }
    // synchronous top-up until `need` words (from chunk 0) are in the ring
    __device__ __forceinline__ void ensure(uint32_t need)
    {
        while (fetched < need) {
            store_chunk(fetched, load_chunk(fetched));
            fetched += 4;
        }
    }
};

// decoder state of one lane: a 96-bit window w0:w1:w2 over the stream
struct LaneDec {
    uint32_t w0, w1, w2;
    uint32_t bit;           // bits of w0 already consumed (< 32 between steps)
    uint32_t wpos;          // position of w0 (words from the wave's chunk 0)
    uint32_t acc;           // running sample (low 16 bits count)
    bool     bad;

    __device__ __forceinline__ void advance(RingFeed &rf)
    {
        if (bit >= 32u) {
            bit -= 32u;
            ++wpos;
            w0 = w1;
            w1 = w2;
            w2 = rf.word(wpos + 2);
        }
    }
    // one code that the table does not hold, through count-leading-zeros (src/deltaRice.c:154-177);
    // `win` = the 32 stream bits at the current position.  Returns the delta.
    __device__ __forceinline__ uint32_t slow_code(uint32_t win, RingFeed &rf, int k, uint32_t kmask, uint32_t &len)
    {
        const uint32_t q = __clz(win);
        uint32_t u;
        if (q >= kEscapeQuotient) {
            bad |= (q > kEscapeQuotient);
            u = (win >> 7) & 0xFFFFu;
            len = kEscapeBits;
        } else {
            len = q + 1 + (uint32_t)k;
            u = (q << k) | ((win >> (32u - len)) & kmask);
        }
        // the periodic refill only covers a block of table-path codes (<= 12 bits each): after an
        // escape make sure the rest of the block (<= 6 more words + the window) is in the ring
        rf.ensure(wpos + 12u);
        const uint32_t h = u >> 1;
        return (u & 1u) ? ~h : h;
    }
    // decodes ONE sample (prologue / epilogue of a wave); bit < 32 on entry and on exit
    __device__ __forceinline__ uint32_t one(RingFeed &rf, uint32_t lut_s, int k, uint32_t kmask)
    {
        const uint32_t win = __funnelshift_l(w1, w0, bit);
        const uint32_t e = lds32(lut_s + ((win >> (32 - kLutBits - 2)) & ((kLutSize - 1) << 2)));
        uint32_t dl, len;
        if (e) {
            dl = (uint32_t)((int32_t)e >> 16);
            len = e & 31u;
        } else {
            dl = slow_code(win, rf, k, kmask, len);
        }
        acc += dl;
        bit += len;
        advance(rf);
        return acc & 0xFFFFu;
    }
    // decodes TWO samples, packed lo | hi << 16; bit < 32 on entry and on exit
    __device__ __forceinline__ uint32_t two(RingFeed &rf, uint32_t lut_s, int k, uint32_t kmask)
    {
        const uint32_t win1 = __funnelshift_l(w1, w0, bit);
        const uint32_t e1 = lds32(lut_s + ((win1 >> (32 - kLutBits - 2)) & ((kLutSize - 1) << 2)));
        uint32_t d1, len1;
        if (e1) {
            d1 = (uint32_t)((int32_t)e1 >> 16);
            len1 = e1 & 31u;
        } else {
            d1 = slow_code(win1, rf, k, kmask, len1);
            bit += len1;
            advance(rf);                 // an escape may cross a word: keep the second window in reach
            len1 = 0;
        }
        const uint32_t b1 = bit + len1;                     // < 44
        const bool hiw = b1 >= 32u;
        const uint32_t win2 = __funnelshift_l(hiw ? w2 : w1, hiw ? w1 : w0, b1);
        const uint32_t e2 = lds32(lut_s + ((win2 >> (32 - kLutBits - 2)) & ((kLutSize - 1) << 2)));
        const uint32_t y1 = acc + d1;
        uint32_t d2, len2;
        if (e2) {
            d2 = (uint32_t)((int32_t)e2 >> 16);
            len2 = e2 & 31u;
        } else {
            bit = b1;
            advance(rf);
            d2 = slow_code(__funnelshift_l(w1, w0, bit), rf, k, kmask, len2);
            bit += len2;
            advance(rf);
            acc = y1 + d2;
            return prmt(y1, acc, 0x5410);
        }
        acc = y1 + d2;
        bit = b1 + len2;                                    // < 56
        advance(rf);
        return prmt(y1, acc, 0x5410);
    }
};

__global__ void __launch_bounds__(kParseMaxWarps * 32, 2) parse_kernel(const ParseParams p)
{
    extern __shared__ __align__(16) uint32_t dsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // shared layout: [pad to 2 KB] one 2 KB ring per warp | table
    const uint32_t dsm_s = (uint32_t)__cvta_generic_to_shared(dsm);
    const uint32_t nwarps = blockDim.x >> 5;
    const uint32_t rings_s = (dsm_s + kRingWords * 128u - 1u) & ~(kRingWords * 128u - 1u);
    const uint32_t lut_s = rings_s + nwarps * (kRingWords * 128u);
    if (lut_s + kLutSize * 4u > dsm_s + p.smem_bytes) {     // launcher and kernel disagree on the layout
        if (threadIdx.x == 0) atomicOr(p.status, kErrStream);
        return;
    }
    const int k = p.k;
    const uint32_t kmask = (1u << k) - 1u;

    for (uint32_t i = threadIdx.x; i < (uint32_t)kLutSize; i += blockDim.x) sts32(lut_s + 4u * i, make_lut_entry(i, k));
    __syncthreads();

    // compressed words are fetched as 16-byte chunks that are aligned in memory: positions are
    // words relative to the aligned address at or below p.comp
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p.comp) >> 2) & 3u);
    const uint32_t *comp_al = p.comp - mis;
    const uint64_t lim_al = p.comp_words + mis;             // end of the stream, aligned-relative
    const uint32_t ngroups = (p.nwaves + 31u) / 32u;

    while (true) {
        uint32_t grp = 0;
        if (lane == 0) grp = atomicAdd(p.ticket, 1u);
        grp = __shfl_sync(0xffffffffu, grp, 0);
        if (grp >= ngroups) break;
        const uint32_t g = grp * 32 + lane;
        const bool active = g < p.nwaves;
        const uint32_t n = active ? __ldg(p.wave_n + g) : 0u;
        const uint64_t rec = active ? __ldg(p.wave_in + g) : 0ull;        // word index of [nwords]
        const uint64_t obase = active ? __ldg(p.wave_out + g) : 0ull;     // sample offset of the wave
        const uint32_t nwords = n ? __ldg(p.comp + rec) : 0u;
        int16_t *optr = p.out + obase;

        RingFeed rf;
        rf.base_al = (rec + 1 + mis) & ~3ull;               // aligned-relative index of the chunk holding the first code word
        rf.gbase = comp_al + rf.base_al;
        rf.comp_al = comp_al;
        rf.lim_al = lim_al;
        rf.mis = mis;
        rf.safe = (rf.base_al >= mis)
            ? (uint32_t)(lim_al - rf.base_al > 0xFFFFFFF0ull ? 0xFFFFFFF0ull : ((lim_al - rf.base_al) & ~3ull)) : 0u;
        rf.ring_b = rings_s + warp * (kRingWords * 128u) + lane * 4u;
        rf.fetched = 0;

        LaneDec d;
        d.wpos = (uint32_t)((rec + 1 + mis) - rf.base_al);
        d.bit = 0; d.acc = 0; d.bad = false; d.w0 = d.w1 = d.w2 = 0;
        __syncwarp();
        if (n) {
#pragma unroll
            for (int c = 0; c < kRingWords / 4; c += 2) {
                const uint4 c0 = rf.load_chunk(4u * c), c1 = rf.load_chunk(4u * c + 4);
                rf.store_chunk(4u * c, c0);
                rf.store_chunk(4u * c + 4, c1);
            }
            rf.fetched = kRingWords;
            d.w0 = rf.word(d.wpos);
            d.w1 = rf.word(d.wpos + 1);
            d.w2 = rf.word(d.wpos + 2);
        }

        // ---- prologue: single samples up to the first 32-byte boundary of the output --------------
        uint32_t left = n;
        {
            const uint32_t a = (uint32_t)((reinterpret_cast<uintptr_t>(optr) >> 1) & 15u);
            uint32_t pro = (16u - a) & 15u;
            if (pro > left) pro = left;
            left -= pro;
            if (pro) rf.ensure(d.wpos + 10u);
            for (; pro; --pro) *optr++ = (int16_t)d.one(rf, lut_s, k, kmask);
        }
        // ---- blocks of 16 samples: 8 steps of two, one 32-byte store (a full sector) -------------------
        uint32_t nblk = left >> 4;
        left &= 15u;
        const uint32_t maxblk = __reduce_max_sync(0xffffffffu, nblk);
        for (uint32_t b = 0; b < maxblk; ++b) {
            if (b < nblk) {
                // ring refill: the chunks requested now go in after the block's steps (a block uses at
                // most 6 words through the table; longer codes top up on demand)
                rf.ensure(d.wpos + 10u);
                uint4 pend0 = make_uint4(0, 0, 0, 0), pend1 = pend0;
                uint32_t pend_at = 0xffffffffu, npend = 0;
                const uint32_t room = (d.wpos & ~3u) + kRingWords - rf.fetched;
                if (room >= 4) {
                    pend_at = rf.fetched;
                    pend0 = rf.load_chunk(rf.fetched);
                    npend = 1;
                    if (room >= 8) {
                        pend1 = rf.load_chunk(rf.fetched + 4);
                        npend = 2;
                    }
                }
                uint4 o0, o1;
                o0.x = d.two(rf, lut_s, k, kmask);
                o0.y = d.two(rf, lut_s, k, kmask);
                o0.z = d.two(rf, lut_s, k, kmask);
                o0.w = d.two(rf, lut_s, k, kmask);
                o1.x = d.two(rf, lut_s, k, kmask);
                o1.y = d.two(rf, lut_s, k, kmask);
                o1.z = d.two(rf, lut_s, k, kmask);
                o1.w = d.two(rf, lut_s, k, kmask);
                stg_256(optr, o0, o1);
                optr += 16;
                if (pend_at == rf.fetched) {                 // (an on-demand top-up may have overtaken it)
                    rf.store_chunk(rf.fetched, pend0);
                    if (npend == 2) rf.store_chunk(rf.fetched + 4, pend1);
                    rf.fetched += 4 * npend;
                }
            }
        }
        // ---- epilogue: the last < 8 samples --------------------------------------------------------
        if (left) rf.ensure(d.wpos + 10u);
        for (; left; --left) *optr++ = (int16_t)d.one(rf, lut_s, k, kmask);

        // the codes must end inside the last word of the record
        if (n) {
            const uint64_t used = (rf.base_al + d.wpos) - (rec + 1 + mis) + (d.bit ? 1u : 0u);
            if (used != nwords) d.bad = true;
        }
        if (d.bad) atomicOr(p.status, kErrStream);
    }
}

int g_dec_sms = 0;

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

int launch_parse_impl(const ParseParams &p, cudaStream_t st)
{
    if (!g_dec_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_dec_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_dec_sms <= 0) g_dec_sms = 148;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        cudaFuncSetAttribute(parse_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr_set = true;
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
    // shared layout: [pad to 2 KB] rings | table.  The dynamic area starts right after the 1 KB the
    // system reserves per CTA (the kernel checks the assumption).
    size_t off = 1024;
    off = (off + kRingWords * 128 - 1) & ~(size_t)(kRingWords * 128 - 1);
    off += (size_t)warps * kRingWords * 128 + (size_t)kLutSize * 4;
    ParseParams pp = p;
    pp.smem_bytes = (uint32_t)(off - 1024);
    if (getenv("DRICE_DEBUG")) {
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, parse_kernel, warps * 32, pp.smem_bytes);
        fprintf(stderr, "parse: grid %u warps %d smem %u occ %d\n", grid, warps, pp.smem_bytes, occ);
    }
    parse_kernel<<<grid, warps * 32, pp.smem_bytes, st>>>(pp);
    return 1;
}

}  // namespace

int launch_locate(const LocateParams &p, cudaStream_t st)
{
    if (p.nchunks == 0) return 0;
    static bool attr_set = false;
    const size_t smem = (size_t)(kLocStages * kLocTileWords + kLocListMax) * sizeof(uint32_t);
    if (!attr_set) {
        cudaFuncSetAttribute(locate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    locate_kernel<<<p.nchunks, kLocThreads, smem, st>>>(p);
    return 1;
}

int launch_parse(const ParseParams &p, int store_bytes, cudaStream_t st)
{
    (void)store_bytes;          // every alignment takes the same path: 16-byte aligned blocks per lane
    if (p.nwaves == 0) return 0;
    if (p.k < 0 || p.k > 15) return -1;
    return launch_parse_impl(p, st);
}

}  // namespace drice
