// drice_kernels.cuh — kernel parameter blocks and launcher prototypes (internal).
//
// Data layout in HBM:
//   raw      int16 samples, all chunks of a batch back to back
//   comp     uint32 words, chunk streams back to back:
//            chunk := [total][record...], record := [nwords][words...]
//   per-batch tables (device): chunk_sample_off[n+1] u64, chunk_wave_off[n+1] u32,
//            chunk_word_off[n+1] u64; per-wave tables for decode: wave_in (u64 word index of
//            the record header), wave_out (u64 sample offset), wave_n (u32 samples)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace drice {

constexpr int      kSamplesPerThread = 16;   // one 32-byte aligned slot per thread
constexpr uint32_t kEscapeQuotient   = 8;    // reference "giveup", src/deltaRice.c:203
constexpr uint32_t kEscapeBits       = 25;   // 8 zeros + 1 + 16 value bits
constexpr int      kEncMaxThreads    = 512;  // CTA size of the long-wave / redo kernels
constexpr int      kEncTileMaxL      = 8192; // longest wave the warp-per-wave kernel takes

// status flags
constexpr uint32_t kErrCapacity = 1u;   // output capacity exceeded
constexpr uint32_t kErrStream   = 2u;   // malformed stream (bad counts / unary run > 8)
constexpr uint32_t kErrTotal    = 4u;   // chunk's own sample count != expected
constexpr uint32_t kErrInternal = 8u;   // launcher and kernel disagree (shared-memory layout): a bug here, not a bad stream

struct EncodeParams {
    const int16_t  *raw;
    uint64_t        raw_samples;        // total samples in `raw` (vector-load guard)
    uint32_t       *out;
    uint64_t        out_cap_words;
    const uint64_t *chunk_sample_off;   // [nchunks+1]
    const uint32_t *chunk_wave_off;     // [nchunks+1]
    uint64_t       *chunk_byte_off;     // [nchunks+1] result (BYTE offsets of the chunk streams)
    uint64_t       *lookback;           // [ntiles] (<= nwaves), zeroed
    uint32_t       *ticket;             // zeroed
    uint32_t       *status;             // error flags (OR-ed)
    uint32_t        nchunks;
    uint32_t        nwaves;
    uint32_t        uniform_wpc;        // >0: every chunk has exactly this many waves
    uint32_t        L;                  // 0 = whole chunk is one wave
    int             k;
};

// what the encode launchers need beyond EncodeParams (kept out of it: the tile kernel's code
// generation is sensitive to the size of its by-value parameter block)
struct EncodeMode {
    // pre-filter mode: 1 = delta (src/deltaRice.c:53-62), 0 = none: the samples are Rice-coded as
    // they are (filter [1], or already filtered by prefilter_kernel)
    int             delta;
    // words of the largest record the context's previous batch of the same shape produced (0 = unknown):
    // sizes the tile kernel's staging; the launch raises *max_words (device) with its own largest record
    uint32_t        words_hint;
    uint32_t       *max_words;
    // scratch of the long-wave encoder (encode_long_scratch_bytes; need not be zeroed), or null
    void           *long_scratch;
    size_t          long_scratch_bytes;
};

// generic pre-filter (src/deltaRice.c:64-74 / :91-102), taps by value
constexpr int kMaxFilter = 16;
struct FilterParams {
    const uint64_t *chunk_sample_off;   // [nchunks+1] (device)
    uint32_t        nchunks;
    uint32_t        L;                  // 0 = whole chunk is one wave
    int             flen;
    int             f[kMaxFilter];
};

struct LocateParams {
    const uint32_t *comp;
    uint64_t        comp_words;         // total words in comp (load guard)
    const uint64_t *chunk_word_off;     // [nchunks+1] (words)
    const uint64_t *chunk_sample_off;   // [nchunks+1]
    const uint32_t *chunk_wave_off;     // [nchunks+1]
    uint64_t       *wave_in;            // [nwaves]
    uint64_t       *wave_out;           // [nwaves]
    uint32_t       *wave_n;             // [nwaves]
    uint32_t       *status;
    uint32_t        nchunks;
    uint32_t        L;                  // 0 = whole chunk
};

struct ParseParams {
    const uint32_t *comp;
    uint64_t        comp_words;
    const uint64_t *wave_in;
    const uint64_t *wave_out;
    const uint32_t *wave_n;
    int16_t        *out;
    uint32_t       *status;
    uint32_t       *ticket;             // zeroed: warp tasks (32 waves each) handed out
    uint32_t        nwaves;
    uint32_t        max_n;              // longest wave in the batch
    uint32_t        smem_bytes;         // dynamic shared memory of the launch (set by the launcher)
    int             k;
    int             identity;           // 1: write the decoded values themselves (no inverse delta)
    unsigned long long *long_state;     // zeroed, parse_long_state_bytes(): chain state of the long-wave parser (or null)
    // lane parser, batches with heavy records: waves sorted by bits per sample (null: waves in order).
    // sort_perm: [nwaves] scratch; sort_counters: 16 zeroed words (histogram + cursors of 8 buckets)
    uint32_t       *sort_perm;
    uint32_t       *sort_counters;
    int             heavy_ok;           // 1: the parser instance whose heavy warps skip the table (batches above k + 3.5 bits per sample)
};

// ---- per-device launch state ------------------------------------------------------------------
// cudaFuncSetAttribute (dynamic shared memory opt-in) and the SM count belong to a DEVICE, and one
// process may hold contexts on several (drice_create(ctx, device)): everything cached about a
// launch is keyed by the current device.
constexpr int kMaxDevices = 64;
inline int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int device_sm_count()
{
    static int sms[kMaxDevices] = {};
    const int dev = current_device();
    if (!sms[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    return sms[dev];
}
// one flag per device: `if (once.first()) cudaFuncSetAttribute(...)`; declare it `static` next to
// the launch it guards (benign race: the attribute calls are idempotent)
struct DeviceOnce {
    unsigned long long done = 0;
    bool first()
    {
        const unsigned long long bit = 1ull << current_device();
        if (done & bit) return false;
        done |= bit;
        return true;
    }
};

#ifdef __CUDACC__
// ---- mbarrier + bulk async copy (TMA 1-D, SASS UBLKCP) ------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    while (!mbar_try_wait(mbar, parity)) {}
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `mbar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
#endif

// launchers (drice_encode.cu / drice_decode.cu); return launches enqueued
int launch_encode(const EncodeParams &p, const EncodeMode &m, uint32_t max_wave_len, cudaStream_t st);
// few waves longer than kEncTileMaxL are encoded by several CTAs per wave: scratch for their segment tables
uint32_t encode_long_maxseg(uint32_t max_wave_len);
size_t encode_long_scratch_bytes(uint32_t nwaves, uint32_t max_wave_len);
int launch_locate(const LocateParams &p, cudaStream_t st);
// header scan instead of the chase (drice_decode.cu): applicability + scratch size, launcher
bool locate_scan_applies(uint32_t L, int k, uint64_t max_chunk_words, size_t nchunks, size_t *scratch_bytes);
int launch_locate_scan(const LocateParams &p, int k, uint64_t max_chunk_words, uint64_t max_chunk_waves, void *scratch, cudaStream_t st);
// very large batches: one warp per chunk walks the chain through global memory (cheaper than reading the whole stream)
bool locate_direct_applies(uint64_t comp_words, uint64_t max_chunk_waves, size_t nchunks);
int launch_locate_direct(const LocateParams &p, int k, cudaStream_t st);
int launch_parse(const ParseParams &p, int store_bytes, cudaStream_t st);
size_t parse_long_state_bytes(uint32_t nwaves, uint32_t max_n);
// out[i] = sum_j in[i-j] * f[j] per wave (mod 2^16); in != out
int launch_prefilter(const FilterParams &p, const int16_t *in, int16_t *out, uint64_t max_chunk_samples, cudaStream_t st);
// in place: y[i] = (d[i] - sum_{j>=1} y[i-j] * f[j]) / f[0] per wave
int launch_postfilter(const FilterParams &p, int16_t *data, uint64_t nwaves_hint, cudaStream_t st);

}  // namespace drice
