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

namespace drice {

namespace {

// ------------------------------------------------------------------------------------
// locate
// ------------------------------------------------------------------------------------
constexpr int kLocThreads   = 128;
constexpr int kLocTileWords = 8192;          // 32 KB per stage, 2 stages

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
__device__ __forceinline__ void locate_load_tile(uint32_t *dst, const uint32_t *comp, int64_t A,
                                                 uint64_t limit)
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
    extern __shared__ __align__(16) uint32_t stile[];   // 2 * kLocTileWords
    __shared__ uint64_t s_cur;
    __shared__ uint32_t s_wave;
    __shared__ int s_done;
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
        s_done = (W == 0);
    }
    __syncthreads();
    if (s_done) return;

    // misalignment of the global address: tiles start at word indices A with (comp + A) 16-byte aligned
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p.comp) >> 2) & 3u);
    auto align_down = [mis](uint64_t w) { return (int64_t)((w + mis) & ~3ull) - (int64_t)mis; };

    int64_t A = align_down(wb + 1);
    int buf = 0;
    locate_load_tile(stile, p.comp, A, we);
    cp_async_commit();
    while (true) {
        // prefetch the sequentially next tile into the other buffer
        const int64_t An = A + kLocTileWords;
        if ((uint64_t)An < we) locate_load_tile(stile + (buf ^ 1) * kLocTileWords, p.comp, An, we);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t *t = stile + buf * kLocTileWords;
            uint64_t cur = s_cur;
            uint32_t w = s_wave;
            const uint64_t tile_end = (uint64_t)(A + kLocTileWords);
            while (w < W && cur < tile_end && cur < we) {
                const uint32_t nw = t[(int64_t)cur - A];
                const uint64_t s0 = (uint64_t)w * Lw;
                p.wave_in[g0 + w] = cur;
                p.wave_out[g0 + w] = sb + s0;
                p.wave_n[g0 + w] = (uint32_t)((total - s0) < Lw ? (total - s0) : Lw);
                cur += (uint64_t)nw + 1;
                ++w;
            }
            s_cur = cur;
            s_wave = w;
            if (w == W) {
                if (cur != we) atomicOr(p.status, kErrStream);
                s_done = 1;
            } else if (cur >= we) {
                atomicOr(p.status, kErrStream);
                // neutralise the waves that could not be located
                for (; w < W; ++w) { p.wave_in[g0 + w] = wb; p.wave_out[g0 + w] = sb; p.wave_n[g0 + w] = 0; }
                s_done = 1;
            }
        }
        __syncthreads();
        if (s_done) break;
        const uint64_t cur = s_cur;
        if (cur < (uint64_t)(An + kLocTileWords)) {
            A = An;                           // the prefetched tile is the one we need
            buf ^= 1;
        } else {                              // long record: jump, drop the prefetch
            cp_async_wait<0>();
            __syncthreads();
            A = align_down(cur);
            locate_load_tile(stile + buf * kLocTileWords, p.comp, A, we);
            cp_async_commit();
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------
// parse
// ------------------------------------------------------------------------------------
constexpr int kParseWarps   = 4;             // warps per CTA
constexpr int kRing         = 16;            // ring words per lane
constexpr int kTile         = 64;            // samples per lane per output tile
constexpr int kTileStrideW  = kTile / 2 + 2; // words per row: 8-byte aligned rows

__device__ __forceinline__ int4 ld_stream_v4(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// loads the 4 words [a, a+4) of comp (a: multiple of 4 relative to a 16-byte aligned address)
__device__ __forceinline__ int4 load_quad(const uint32_t *comp, int64_t a, uint64_t limit)
{
    if (a >= 0 && (uint64_t)a + 4 <= limit) return ld_stream_v4(reinterpret_cast<const int4 *>(comp + a));
    int4 r;
    int *e = &r.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) e[i] = (a + i >= 0 && (uint64_t)(a + i) < limit) ? (int)comp[a + i] : 0;
    return r;
}

template <int K, int STORE_BYTES>
__global__ void __launch_bounds__(kParseWarps * 32) parse_kernel(const ParseParams p)
{
    __shared__ uint32_t s_ring[kParseWarps][kRing][32];
    __shared__ __align__(16) uint32_t s_tile[kParseWarps][32][kTileStrideW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = (blockIdx.x * kParseWarps + warp) * 32 + lane;
    uint32_t(*ring)[32] = s_ring[warp];
    uint32_t(*tile)[kTileStrideW] = s_tile[warp];

    const bool active = g < p.nwaves;
    const uint32_t n = active ? p.wave_n[g] : 0u;
    const uint64_t rec = active ? p.wave_in[g] : 0ull;        // word index of [nwords]
    const uint64_t obase = active ? p.wave_out[g] : 0ull;     // sample offset of the wave
    uint32_t nmax = n;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, d));
    if (nmax == 0) return;

    // ring addressing uses word indices relative to a 16-byte aligned origin
    const int mis = (int)((reinterpret_cast<uintptr_t>(p.comp) >> 2) & 3u);
    const uint32_t nwords = n ? p.comp[rec] : 0u;
    int64_t wi = (int64_t)rec + 1;                             // current word (absolute index)
    int64_t loaded = ((wi + mis) & ~3ll) - mis;                // ring holds [.., loaded)
    // prime the ring with 3 quads
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int4 qd = load_quad(p.comp, loaded, p.comp_words);
        const uint32_t s = (uint32_t)(loaded + mis);
        ring[(s + 0) & (kRing - 1)][lane] = qd.x;
        ring[(s + 1) & (kRing - 1)][lane] = qd.y;
        ring[(s + 2) & (kRing - 1)][lane] = qd.z;
        ring[(s + 3) & (kRing - 1)][lane] = qd.w;
        loaded += 4;
    }
    uint32_t w0 = ring[(uint32_t)(wi + mis) & (kRing - 1)][lane];
    uint32_t w1 = ring[(uint32_t)(wi + mis + 1) & (kRing - 1)][lane];
    uint32_t bit = 0;
    int4 pend = make_int4(0, 0, 0, 0);
    bool pend_valid = false;
    uint32_t acc = 0;
    bool bad = false;

    for (uint32_t t0 = 0; t0 < nmax; t0 += kTile) {
#pragma unroll 1
        for (int jj = 0; jj < kTile; jj += 4) {
            // ---- ring maintenance: commit last group's load, issue the next ----------------
            if (pend_valid) {
                const uint32_t s = (uint32_t)(loaded + mis);
                ring[(s + 0) & (kRing - 1)][lane] = pend.x;
                ring[(s + 1) & (kRing - 1)][lane] = pend.y;
                ring[(s + 2) & (kRing - 1)][lane] = pend.z;
                ring[(s + 3) & (kRing - 1)][lane] = pend.w;
                loaded += 4;
                pend_valid = false;
            }
            if (t0 + jj < n && loaded - wi <= kRing - 4) {
                pend = load_quad(p.comp, loaded, p.comp_words);
                pend_valid = true;
            }
            uint32_t pk[2];
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
                const uint32_t j = t0 + jj + u4;
                if (j < n) {
                    const uint32_t win = __funnelshift_l(w1, w0, bit);
                    const uint32_t q = __clz(win);
                    uint32_t u, len;
                    if (q >= kEscapeQuotient) {
                        bad |= (q > kEscapeQuotient);
                        u = (win >> 7) & 0xFFFFu;
                        len = kEscapeBits;
                    } else {
                        len = q + (K + 1);
                        u = (q << K) | ((win >> (32u - len)) & ((1u << K) - 1u));
                    }
                    const uint32_t h = u >> 1;
                    acc += (u & 1u) ? ~h : h;                  // un-zig-zag (:172-177) + running sum (:84-89)
                    bit += len;
                    if (bit >= 32u) {
                        bit -= 32u;
                        ++wi;
                        w0 = w1;
                        w1 = ring[(uint32_t)(wi + mis + 1) & (kRing - 1)][lane];
                    }
                }
                if (u4 & 1) pk[u4 >> 1] |= acc << 16; else pk[u4 >> 1] = acc & 0xFFFFu;
            }
            *reinterpret_cast<uint2 *>(&tile[lane][jj >> 1]) = make_uint2(pk[0], pk[1]);
        }
        __syncwarp();
        // ---- store the 32 x kTile tile: row r = wave of lane r --------------------------
        if (STORE_BYTES == 8) {
            // 16 lanes x 8 bytes per row, two rows per instruction
            const int half = lane >> 4, col = (lane & 15) * 4;       // col in samples
#pragma unroll 4
            for (int it = 0; it < 16; ++it) {
                const int r = it * 2 + half;
                const uint32_t rn = __shfl_sync(0xffffffffu, n, r);
                const uint64_t ro = __shfl_sync(0xffffffffu, obase, r);
                const uint32_t cnt = rn > t0 ? min(rn - t0, (uint32_t)kTile) : 0u;
                const uint2 v = *reinterpret_cast<const uint2 *>(&tile[r][col >> 1]);
                int16_t *dst = p.out + ro + t0 + col;
                if ((uint32_t)col + 4 <= cnt) {
                    *reinterpret_cast<uint2 *>(dst) = v;
                } else {
                    const uint32_t e[2] = {v.x, v.y};
                    for (int s = 0; s < 4; ++s)
                        if ((uint32_t)(col + s) < cnt) dst[s] = (int16_t)(e[s >> 1] >> ((s & 1) * 16));
                }
            }
        } else {
            // generic alignment: 2-byte stores, one row per instruction, 2 samples per lane
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                const uint32_t rn = __shfl_sync(0xffffffffu, n, r);
                const uint64_t ro = __shfl_sync(0xffffffffu, obase, r);
                const uint32_t cnt = rn > t0 ? min(rn - t0, (uint32_t)kTile) : 0u;
                int16_t *dst = p.out + ro + t0;
                for (int s = lane; s < (int)cnt; s += 32) {
                    const uint32_t wv = tile[r][s >> 1];
                    dst[s] = (int16_t)(wv >> ((s & 1) * 16));
                }
            }
        }
        __syncwarp();
    }
    // the codes must end inside the last word of the record
    if (n) {
        const uint64_t used = (uint64_t)(wi - (int64_t)(rec + 1)) + (bit ? 1u : 0u);
        if (used != nwords) bad = true;
    }
    if (bad) atomicOr(p.status, kErrStream);
}

template <int K>
int launch_parse_k(const ParseParams &p, int store_bytes, cudaStream_t st)
{
    const uint32_t per_cta = kParseWarps * 32;
    const uint32_t grid = (p.nwaves + per_cta - 1) / per_cta;
    if (store_bytes >= 8)
        parse_kernel<K, 8><<<grid, per_cta, 0, st>>>(p);
    else
        parse_kernel<K, 2><<<grid, per_cta, 0, st>>>(p);
    return 1;
}

}  // namespace

int launch_locate(const LocateParams &p, cudaStream_t st)
{
    if (p.nchunks == 0) return 0;
    static bool attr_set = false;
    const size_t smem = 2 * kLocTileWords * sizeof(uint32_t);
    if (!attr_set) {
        cudaFuncSetAttribute(locate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    locate_kernel<<<p.nchunks, kLocThreads, smem, st>>>(p);
    return 1;
}

int launch_parse(const ParseParams &p, int store_bytes, cudaStream_t st)
{
    if (p.nwaves == 0) return 0;
    switch (p.k) {
#define DRICE_CASE(K) case K: return launch_parse_k<K>(p, store_bytes, st);
        DRICE_CASE(0) DRICE_CASE(1) DRICE_CASE(2) DRICE_CASE(3) DRICE_CASE(4) DRICE_CASE(5)
        DRICE_CASE(6) DRICE_CASE(7) DRICE_CASE(8) DRICE_CASE(9) DRICE_CASE(10) DRICE_CASE(11)
        DRICE_CASE(12) DRICE_CASE(13) DRICE_CASE(14) DRICE_CASE(15)
#undef DRICE_CASE
    }
    return -1;
}

}  // namespace drice
