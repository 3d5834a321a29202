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

#include <cstdlib>

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
// Rice parsing is a serial chain per wave (a code's length is only known once its unary
// prefix has been read), so the parallelism is across waves: one LANE per wave, a warp takes
// 32 consecutive waves per ticket.  What makes a lane fast is how many instructions it spends
// per sample:
//   * multi-symbol table: the next 12 stream bits index a shared-memory table built for the
//     launch's k (4096 x 8 bytes) whose entry holds up to THREE complete codes already turned
//     into running delta sums, the bits they consume and their count, so one lookup + two
//     packed adds yields up to three samples.  Entries with no complete code (escape, or a code
//     longer than 12 bits) send the lane through a count-leading-zeros path for one sample;
//   * the compressed words reach the lane through a private 32-word ring in shared memory
//     filled by 16-byte cp.async (no registers, latency hidden one refill period ahead);
//   * samples are written to a per-warp shared tile (32 lanes x 32 samples) that the warp then
//     stores to HBM row by row with 16-byte (or 8 / 2 byte, by alignment) coalesced stores.
constexpr int kLutBits     = 12;
constexpr int kLutSize     = 1 << kLutBits;
constexpr int kTS          = 32;                 // samples per lane per output tile
constexpr int kRowW        = 18;                 // words per tile row: 32 samples + 3 overshoot, 8-byte aligned
constexpr int kRingWords   = 32;                 // ring words per lane (8 chunks of 16 bytes)
constexpr int kWarpSmemW   = kRingWords * 32 + 32 * kRowW + 32 * 2 + 32;   // ring | tile | row base (u64) | row n
constexpr int kRefillEvery = 16;                 // lookups between ring refills (<= 12.5 words consumed)

__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, uint32_t src_bytes)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// table entry for the 12 bits `idx` (MSB = next stream bit), Rice parameter 2^k:
//   x = S1 | S2 << 16,  y = S3 | nbits << 16 | 2*count << 24
// S_i = sum of the first i decoded deltas (src/deltaRice.c:161-177), the unused ones repeat the
// last, so S3 is always the total.  count = 0: no complete non-escape code in the window.
__device__ __forceinline__ uint2 make_lut_entry(uint32_t idx, int k)
{
    const uint32_t bits = idx << (32 - kLutBits);
    uint32_t pos = 0, cnt = 0;
    int sum = 0;
    int S[3] = {0, 0, 0};
    while (cnt < 3) {
        const uint32_t rem = kLutBits - pos;
        const uint32_t win = bits << pos;
        const uint32_t q = win ? (uint32_t)__clz(win) : 32u;
        if (q >= rem || q >= kEscapeQuotient) break;
        const uint32_t len = q + 1 + (uint32_t)k;
        if (len > rem) break;
        const uint32_t r = (win >> (32 - len)) & ((1u << k) - 1u);
        const uint32_t u = (q << k) | r;
        sum += (u & 1u) ? -(int)((u + 1) >> 1) : (int)(u >> 1);
        S[cnt++] = sum;
        pos += len;
    }
    for (uint32_t i = cnt; i < 3 && cnt; ++i) S[i] = S[cnt - 1];
    uint2 e;
    e.x = ((uint32_t)S[0] & 0xFFFFu) | ((uint32_t)S[1] << 16);
    e.y = ((uint32_t)S[2] & 0xFFFFu) | (pos << 16) | ((2u * cnt) << 24);
    return e;
}

template <int STORE_BYTES>
__global__ void __launch_bounds__(1024, 1) parse_kernel(const ParseParams p)
{
    extern __shared__ __align__(16) uint32_t dsm[];
    uint2 *lut = reinterpret_cast<uint2 *>(dsm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wsm = dsm + 2 * kLutSize + (size_t)warp * kWarpSmemW;
    uint32_t *ring = wsm;                                   // [chunk 0..7][lane][4 words]
    uint32_t *tile = ring + kRingWords * 32;                // [row = lane][kRowW]
    uint64_t *s_rowbase = reinterpret_cast<uint64_t *>(tile + 32 * kRowW);
    uint32_t *s_rown = reinterpret_cast<uint32_t *>(s_rowbase + 32);
    const int k = p.k;

    for (uint32_t i = threadIdx.x; i < (uint32_t)kLutSize; i += blockDim.x) lut[i] = make_lut_entry(i, k);
    __syncthreads();

    // the ring holds 16-byte chunks that are aligned in memory: positions are words relative to
    // the aligned address at or below p.comp
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p.comp) >> 2) & 3u);
    const uint32_t *comp_al = p.comp - mis;
    const uint64_t lim_al = p.comp_words + mis;             // end of the stream, aligned-relative
    uint32_t *ring_lane = ring + lane * 4;
    unsigned char *row_bytes = reinterpret_cast<unsigned char *>(tile + lane * kRowW);
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
        const uint32_t nmax = __reduce_max_sync(0xffffffffu, n);
        if (nmax == 0) continue;
        const uint32_t nwords = n ? __ldg(p.comp + rec) : 0u;

        // ---- ring: chunk `c` of this lane sits at ring_lane + (c & 7) * 128 words -----------
        const uint64_t base_al = (rec + 1 + mis) & ~3ull;   // aligned-relative index of the chunk holding the first code word
        uint32_t wpos = (uint32_t)((rec + 1 + mis) - base_al);   // position of w0, words from base_al
        uint32_t fetched = 0;                               // words requested so far (multiple of 4), from base_al
        auto issue_chunk = [&](uint32_t at) {
            uint32_t *dst = ring_lane + ((at >> 2) & 7u) * 128u;
            const uint64_t aw = base_al + at;               // aligned-relative word index of the chunk
            if (aw >= mis && aw + 4 <= lim_al) {
                cp_async16_zfill(dst, comp_al + aw, 16u);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) dst[e] = (aw + e >= mis && aw + e < lim_al) ? comp_al[aw + e] : 0u;
            }
        };
        auto ring_word = [&](uint32_t x) -> uint32_t {
            return ring_lane[((x & 28u) << 5) + (x & 3u)];
        };
        __syncwarp();
        if (n) {
#pragma unroll
            for (int c = 0; c < 8; ++c) issue_chunk(4u * c);
            fetched = 32;
        }
        s_rowbase[lane] = obase;
        s_rown[lane] = n;
        cp_async_wait_all();
        __syncwarp();
        uint32_t w0 = ring_word(wpos), w1 = ring_word(wpos + 1);
        uint32_t bit = 0;
        uint32_t acc2 = 0;                                  // running sample in both halves
        uint32_t jb = 0;                                    // bytes of samples in the current tile row
        bool bad = false;

        for (uint32_t t0 = 0; t0 < nmax; t0 += kTS) {
            const int32_t left = (int32_t)n - (int32_t)t0;  // samples of this lane's wave from t0 on
            const uint32_t limit_b = left >= kTS ? 2u * kTS : (left > 0 ? 2u * (uint32_t)left : 0u);
            const bool exact = left <= kTS + 2;             // end of the wave: one code at a time, exact end position
            uint32_t it = 0;
            while (jb < limit_b) {
                if ((it++ & (kRefillEvery - 1)) == 0) {
                    // everything requested earlier has landed; top the ring up (the chunk holding
                    // w0 must stay)
                    cp_async_wait_all();
                    const uint32_t room_end = (wpos & ~3u) + kRingWords;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (fetched + 4 <= room_end) {
                            issue_chunk(fetched);
                            fetched += 4;
                        }
                    }
                }
                const uint32_t win = __funnelshift_l(w1, w0, bit);
                const uint2 e = lut[win >> (32 - kLutBits)];
                const uint32_t cnt2 = e.y >> 24;
                if (cnt2 != 0 && !exact) {
                    const uint32_t y12 = __vadd2(e.x, acc2);
                    const uint32_t y3 = __vadd2(e.y, acc2);
                    unsigned char *q = row_bytes + jb;
                    *reinterpret_cast<uint16_t *>(q) = (uint16_t)y12;
                    *reinterpret_cast<uint16_t *>(q + 2) = (uint16_t)(y12 >> 16);
                    *reinterpret_cast<uint16_t *>(q + 4) = (uint16_t)y3;
                    acc2 = prmt(y3, 0, 0x1010);
                    jb += cnt2;
                    bit += prmt(e.y, 0, 0x4442);
                } else {
                    // one code through count-leading-zeros (src/deltaRice.c:154-177)
                    const uint32_t q = __clz(win);
                    uint32_t u, len;
                    if (q >= kEscapeQuotient) {
                        bad |= (q > kEscapeQuotient);
                        u = (win >> 7) & 0xFFFFu;
                        len = kEscapeBits;
                    } else {
                        len = q + 1 + (uint32_t)k;
                        u = (q << k) | ((win >> (32u - len)) & ((1u << k) - 1u));
                    }
                    const uint32_t h = u >> 1;
                    const uint32_t y = (acc2 + ((u & 1u) ? ~h : h)) & 0xFFFFu;
                    *reinterpret_cast<uint16_t *>(row_bytes + jb) = (uint16_t)y;
                    acc2 = y | (y << 16);
                    jb += 2;
                    bit += len;
                }
                if (bit >= 32u) {
                    bit -= 32u;
                    ++wpos;
                    w0 = w1;
                    w1 = ring_word(wpos + 1);
                }
            }
            __syncwarp();
            // ---- store the tile: row r = wave of lane r, samples [t0, t0 + 32) -----------------
            if (STORE_BYTES == 16) {
                const int sub = lane >> 2, col = (lane & 3) * 8;              // 4 lanes x 16 bytes per row
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) {
                    const int r = itr * 8 + sub;
                    const uint32_t rn = s_rown[r];
                    const uint32_t cntr = rn > t0 ? min(rn - t0, (uint32_t)kTS) : 0u;
                    const uint2 a = *reinterpret_cast<const uint2 *>(tile + r * kRowW + (col >> 1));
                    const uint2 b = *reinterpret_cast<const uint2 *>(tile + r * kRowW + (col >> 1) + 2);
                    int16_t *dst = p.out + s_rowbase[r] + t0 + col;
                    if ((uint32_t)col + 8 <= cntr) {
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(a.x, a.y, b.x, b.y);
                    } else {
                        const uint32_t ev[4] = {a.x, a.y, b.x, b.y};
                        for (int s = 0; s < 8; ++s)
                            if ((uint32_t)(col + s) < cntr) dst[s] = (int16_t)(ev[s >> 1] >> ((s & 1) * 16));
                    }
                }
            } else if (STORE_BYTES == 8) {
                const int sub = lane >> 3, col = (lane & 7) * 4;              // 8 lanes x 8 bytes per row
#pragma unroll
                for (int itr = 0; itr < 8; ++itr) {
                    const int r = itr * 4 + sub;
                    const uint32_t rn = s_rown[r];
                    const uint32_t cntr = rn > t0 ? min(rn - t0, (uint32_t)kTS) : 0u;
                    const uint2 a = *reinterpret_cast<const uint2 *>(tile + r * kRowW + (col >> 1));
                    int16_t *dst = p.out + s_rowbase[r] + t0 + col;
                    if ((uint32_t)col + 4 <= cntr) {
                        *reinterpret_cast<uint2 *>(dst) = a;
                    } else {
                        const uint32_t ev[2] = {a.x, a.y};
                        for (int s = 0; s < 4; ++s)
                            if ((uint32_t)(col + s) < cntr) dst[s] = (int16_t)(ev[s >> 1] >> ((s & 1) * 16));
                    }
                }
            } else {
                // generic alignment: 2-byte stores, one row per instruction
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    const uint32_t rn = s_rown[r];
                    const uint32_t cntr = rn > t0 ? min(rn - t0, (uint32_t)kTS) : 0u;
                    int16_t *dst = p.out + s_rowbase[r] + t0;
                    if ((uint32_t)lane < cntr) {
                        const uint32_t wv = tile[r * kRowW + (lane >> 1)];
                        dst[lane] = (int16_t)(wv >> ((lane & 1) * 16));
                    }
                }
            }
            __syncwarp();
            // samples decoded past the tile's end open the next tile
            if (jb > 2u * kTS) {
                const uint32_t over = jb - 2u * kTS;             // 2 or 4 bytes
                const uint32_t c0 = *reinterpret_cast<const uint32_t *>(row_bytes + 2 * kTS);
                *reinterpret_cast<uint32_t *>(row_bytes) = c0;
                jb = over;
            } else {
                jb = 0;
            }
        }
        // the codes must end inside the last word of the record
        if (n) {
            const uint64_t used = (base_al + wpos) - (rec + 1 + mis) + (bit ? 1u : 0u);
            if (used != nwords) bad = true;
        }
        if (bad) atomicOr(p.status, kErrStream);
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

template <int STORE_BYTES>
int launch_parse_t(const ParseParams &p, cudaStream_t st)
{
    if (!g_dec_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_dec_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_dec_sms <= 0) g_dec_sms = 148;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(parse_kernel<STORE_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    const uint32_t ngroups = (p.nwaves + 31u) / 32u;
    static int max_warps = 0;
    if (!max_warps) {
        const char *e = getenv("DRICE_DEC_WARPS");
        max_warps = e ? atoi(e) : 24;
        if (max_warps < 1) max_warps = 1;
        if (max_warps > 28) max_warps = 28;
    }
    int warps = pick_parse_warps(ngroups, g_dec_sms, max_warps);
    uint32_t grid = (uint32_t)g_dec_sms;
    if ((uint64_t)grid * warps > ngroups) {
        // small batch: spread the groups over the SMs
        grid = (ngroups + warps - 1) / warps;
        if (grid < (uint32_t)g_dec_sms && ngroups >= (uint32_t)g_dec_sms) grid = (uint32_t)g_dec_sms;
        if (grid > (uint32_t)g_dec_sms) grid = (uint32_t)g_dec_sms;
        warps = (int)((ngroups + grid - 1) / grid);
        if (warps < 1) warps = 1;
    }
    const size_t smem = (size_t)(2 * kLutSize + warps * kWarpSmemW) * sizeof(uint32_t);
    parse_kernel<STORE_BYTES><<<grid, warps * 32, smem, st>>>(p);
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
    if (p.k < 0 || p.k > 15) return -1;
    if (store_bytes >= 16) return launch_parse_t<16>(p, st);
    if (store_bytes >= 8) return launch_parse_t<8>(p, st);
    return launch_parse_t<2>(p, st);
}

}  // namespace drice
