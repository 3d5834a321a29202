// drice_api.cu — C-ABI of include/deltarice_b200.h: parameter parsing, per-batch tables,
// the device entry points and the chunk scheduler (host-pointer entry points).
//
// Host-side counterpart of the reference's chunk layer (paths relative to /root/reference):
//   parseCD_VALUES                     src/deltaRice.c:248-291
//   determinePowerOf2                  src/deltaRice.c:114-136
//   writeWholeCompressedByteString     src/deltaRice.c:383-465  (per chunk -> per batch here)
//   readWholeCompressedByteString      src/deltaRice.c:301-358
#include "../../include/deltarice_b200.h"
#include "drice_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace drice;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void  *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Slot {                 // one in-flight sub-batch of the chunk scheduler
    DevBuf raw, comp, offs;   // device raw samples, device stream, device chunk byte offsets (+status)
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;
    std::vector<uint64_t> h_offs;
    uint64_t *h_offs_pinned = nullptr;
    size_t h_offs_cap = 0;
    size_t c0 = 0, c1 = 0;    // chunk range
    bool busy = false;
};

}  // namespace

struct drice_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;       // default work stream
    cudaStream_t s_in = nullptr, s_out = nullptr;   // the device's shared copy streams (not owned)
    std::string err;
    uint64_t launches = 0;
    // pre-filter of the next calls (drice_set_filter): 0 = delta [1,-1] (fused), 1 = none ([1]), 2 = generic taps
    int filter_mode = 0;
    int filter_len = 2;
    int filter[drice::kMaxFilter] = {1, -1};
    DevBuf d_filt;                       // generic filter: pre-filtered samples of the batch
    DevBuf d_scan;                       // locate by scanning: per-tile header candidates
    // the largest record (words) of the previous batch sizes the encoder's per-wave staging of the
    // next one.  The kernel raises d_hint[0]; the next call's prep kernel moves it to h_hint (mapped
    // pinned memory, read by the host without a synchronisation) and clears it.
    uint32_t *d_hint = nullptr;
    volatile uint32_t *h_hint = nullptr;
    uint32_t hint_wave = 0;              // wave length / Rice parameter the hint was measured on
    int hint_k = -1;
    // pageable caller buffers: two pinned pieces the copies are staged through (see staged_h2d)
    char *h_piece[2] = {nullptr, nullptr};
    cudaEvent_t ev_piece[2] = {nullptr, nullptr};

    // optional per-kernel timing (drice_timing_*): event pairs recorded around launches
    bool timing = false;
    struct TimedLaunch { int kind; cudaEvent_t e0, e1; };
    std::vector<TimedLaunch> timed;             // pending, not yet resolved
    std::vector<cudaEvent_t> ev_pool;
    double   t_ms[DRICE_NUM_KERNELS] = {0, 0, 0};
    uint64_t t_n[DRICE_NUM_KERNELS] = {0, 0, 0};

    // per-call tables (host pinned staging + device)
    void  *h_tab = nullptr;
    size_t h_tab_cap = 0;
    DevBuf d_tab;                        // chunk tables
    DevBuf d_scratch;                    // look-back words / wave tables, ticket, status
    cudaEvent_t ev_tab = nullptr;        // tables uploaded (h_tab reusable)
    // the scratch (tickets, look-back words, wave tables, scan candidates) is one set per context: a call
    // enqueued on ANOTHER stream than the previous one first waits for that call's last kernel
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    bool ev_tab_pending = false;

    // synchronous-call helpers
    DevBuf d_offs;                       // chunk byte offsets + status for the *_dev sync forms
    uint64_t *h_sync = nullptr;          // pinned landing zone
    size_t h_sync_cap = 0;

    Slot slots[3];
};

namespace {

int fail(drice_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}
int cuda_fail(drice_ctx *ctx, cudaError_t e, const char *what)
{
    return fail(ctx, DRICE_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define DR_CUDA(ctx, call)                                         \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

size_t wave_bound_words(size_t n) { return (25 * n + 31) / 32; }

// geometry of a batch derived from the caller's chunk sizes
struct Geometry {
    std::vector<uint32_t> wave_off;    // [nchunks+1]
    uint32_t nwaves = 0;
    uint32_t uniform_wpc = 0;
    uint32_t max_wave = 0;
    uint32_t Lk = 0;                   // kernel L (0 = whole chunk)
    bool     wave_offsets_mult4 = true;
    uint32_t align_samples = 8;        // largest power of two <= 8 dividing every wave start and end
};

int build_geometry(drice_ctx *ctx, const uint64_t *off, size_t nchunks, int64_t L, bool encode, Geometry &g)
{
    if (L == 0 || L < -1) return fail(ctx, DRICE_E_PARAM, "WaveformLength must be -1 or >= 1");
    if (L > 0x7fffffffll) return fail(ctx, DRICE_E_PARAM, "WaveformLength too large");
    if (nchunks > 0x7ffffffeull) return fail(ctx, DRICE_E_PARAM, "too many chunks");
    g.Lk = L < 0 ? 0u : (uint32_t)L;
    g.wave_off.resize(nchunks + 1);
    uint64_t nw = 0;
    bool uniform = true;
    uint64_t first_w = 0;
    uint64_t align_bits = off[nchunks] | 8;
    for (size_t c = 0; c < nchunks; ++c) {
        if (off[c + 1] < off[c]) return fail(ctx, DRICE_E_PARAM, "chunk_sample_off must be non-decreasing");
        const uint64_t total = off[c + 1] - off[c];
        // per-chunk counts are 32-bit ints in the format (src/deltaRice.c:306,:389)
        if (total > 0x7fffffffull) return fail(ctx, DRICE_E_PARAM, "chunk larger than 2^31-1 samples");
        const uint64_t Lw = g.Lk ? g.Lk : total;
        uint64_t w = total ? (total + Lw - 1) / Lw : 0;
        if (encode && total == 0) w = 1;      // pseudo wave: emits the [0] chunk header
        g.wave_off[c] = (uint32_t)nw;
        // every chunk but the last must hold the same number of waves for wave -> chunk by division
        if (c == 0) first_w = w; else if (w != first_w && !(c + 1 == nchunks && w <= first_w)) uniform = false;
        align_bits |= off[c];
        if (w > 1) align_bits |= Lw;
        nw += w;
        if (nw > 0xfffffff0ull) return fail(ctx, DRICE_E_PARAM, "too many waves in one batch");
        g.max_wave = std::max<uint32_t>(g.max_wave, (uint32_t)std::min<uint64_t>(Lw, total));
        if ((off[c] & 3) || (w > 1 && (Lw & 3))) g.wave_offsets_mult4 = false;
    }
    g.wave_off[nchunks] = (uint32_t)nw;
    g.nwaves = (uint32_t)nw;
    g.uniform_wpc = (uniform && nchunks > 0 && first_w > 0) ? (uint32_t)first_w : 0u;
    g.align_samples = (uint32_t)(align_bits & (~align_bits + 1));
    return DRICE_OK;
}

// Small tables and results (offsets, status words) cross PCIe through a one-CTA kernel that reads /
// writes MAPPED pinned host memory, not through the copy engines: a copy engine is a FIFO shared by
// every stream of the process, so a 100-byte table queued behind another handle's 64 MB sub-batch
// would stall its whole pipeline (seen when encode and decode batches run side by side).
__global__ void word_copy_kernel(uint32_t *dst, const uint32_t *src, uint32_t nwords)
{
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
}
// bytes must be a multiple of 4; `dst` / `src` are device pointers or cudaMallocHost pointers
int small_copy(drice_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return DRICE_OK;
    const uint32_t nwords = (uint32_t)(bytes / 4);
    word_copy_kernel<<<1, nwords < 1024u ? ((nwords + 31u) & ~31u) : 1024u, 0, st>>>((uint32_t *)dst, (const uint32_t *)src, nwords);
    DR_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return DRICE_OK;
}

// one launch prepares a call: the tables come over from mapped host memory, the scratch range the
// kernels expect zeroed (tickets, look-back words) is cleared, the status word reset
__global__ void prep_kernel(uint32_t *dst, const uint32_t *src, uint32_t nwords, uint4 *zero, size_t zero_vec, uint32_t *status,
                            uint32_t *hint_dev, volatile uint32_t *hint_host)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < nwords; i += nth) dst[i] = src[i];
    for (size_t i = tid; i < zero_vec; i += nth) zero[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0 && status) *status = 0;
    if (tid == 0 && hint_dev) {                      // the previous batch's largest record -> host
        const uint32_t h = hint_dev[0];
        if (h) { hint_host[0] = h; hint_dev[0] = 0; }
    }
}

// uploads [chunk_sample_off u64 (n+1)] [second u64 table (n+1), optional] [wave_off u32 (n+1)], zeroes
// `zero_bytes` (rounded up to 16) at `zero_p` (16-byte aligned, may be null) and resets *status
int upload_tables(drice_ctx *ctx, const uint64_t *t0, const uint64_t *t1, const uint32_t *wave_off,
                  size_t nchunks, cudaStream_t st, uint64_t **d_t0, uint64_t **d_t1, uint32_t **d_w,
                  void *zero_p, size_t zero_bytes, uint32_t *status, bool with_hint = false)
{
    const size_t n1 = nchunks + 1;
    const size_t bytes = n1 * 8 * 2 + n1 * 4;
    // (first device work of every call) a different stream than the previous call's: order behind it
    if (ctx->have_last && ctx->last_stream != st) DR_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_last, 0));
    if (ctx->ev_tab_pending) {
        DR_CUDA(ctx, cudaEventSynchronize(ctx->ev_tab));
        ctx->ev_tab_pending = false;
    }
    if (bytes > ctx->h_tab_cap) {
        if (ctx->h_tab) cudaFreeHost(ctx->h_tab);
        ctx->h_tab = nullptr;
        ctx->h_tab_cap = 0;
        DR_CUDA(ctx, cudaMallocHost(&ctx->h_tab, bytes * 2));
        ctx->h_tab_cap = bytes * 2;
    }
    // the previous call's kernels may still read d_tab: same-stream ordering covers `st`;
    // growing the buffer frees it, so drain first
    if (bytes > ctx->d_tab.cap) DR_CUDA(ctx, cudaDeviceSynchronize());
    DR_CUDA(ctx, ctx->d_tab.reserve(bytes));
    char *h = (char *)ctx->h_tab;
    memcpy(h, t0, n1 * 8);
    if (t1) memcpy(h + n1 * 8, t1, n1 * 8); else memset(h + n1 * 8, 0, n1 * 8);
    memcpy(h + n1 * 16, wave_off, n1 * 4);
    {
        const size_t zero_vec = zero_p ? (zero_bytes + 15) / 16 : 0;
        size_t work = std::max<size_t>(bytes / 4, zero_vec);
        unsigned grid = (unsigned)std::min<size_t>((work + 255) / 256, 296);
        if (grid < 1) grid = 1;
        prep_kernel<<<grid, 256, 0, st>>>((uint32_t *)ctx->d_tab.p, (const uint32_t *)h, (uint32_t)(bytes / 4), (uint4 *)zero_p, zero_vec, status,
                                          with_hint ? ctx->d_hint : nullptr, with_hint ? ctx->h_hint : nullptr);
        DR_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    DR_CUDA(ctx, cudaEventRecord(ctx->ev_tab, st));
    ctx->ev_tab_pending = true;
    *d_t0 = (uint64_t *)ctx->d_tab.p;
    *d_t1 = (uint64_t *)((char *)ctx->d_tab.p + n1 * 8);
    *d_w = (uint32_t *)((char *)ctx->d_tab.p + n1 * 16);
    return DRICE_OK;
}

cudaEvent_t timing_event(drice_ctx *ctx)
{
    if (!ctx->ev_pool.empty()) {
        cudaEvent_t e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// brackets one kernel launch with events on its stream when timing is on
struct TimedScope {
    drice_ctx *ctx;
    cudaStream_t st;
    int kind;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    TimedScope(drice_ctx *c, int k, cudaStream_t s) : ctx(c), st(s), kind(k)
    {
        if (!ctx->timing) return;
        e0 = timing_event(ctx);
        e1 = timing_event(ctx);
        if (e0) cudaEventRecord(e0, st);
    }
    ~TimedScope()
    {
        if (!ctx->timing) return;
        if (e1) cudaEventRecord(e1, st);
        if (e0 && e1) ctx->timed.push_back({kind, e0, e1});
    }
};

int status_to_error(drice_ctx *ctx, uint32_t status)
{
    if (status == 0) return DRICE_OK;
    if (status & kErrInternal)
        return fail(ctx, DRICE_E_CUDA, "internal error: the parser's shared-memory layout does not match its launch (not a stream problem)");
    if (status & kErrCapacity) return fail(ctx, DRICE_E_CAPACITY, "output buffer too small for the compressed batch");
    if (status & kErrTotal) return fail(ctx, DRICE_E_STREAM, "chunk stream's sample count differs from the expected chunk size");
    return fail(ctx, DRICE_E_STREAM, "malformed Delta-Rice stream");
}

}  // namespace

// ======================================================================================
// parameter helpers
// ======================================================================================
extern "C" int drice_abi_version(void) { return DRICE_ABI_VERSION; }

extern "C" int drice_log2_param(int M)
{
    if (M <= 0 || (M & (M - 1)) != 0) return -1;
    int k = 0;
    while ((1 << k) != M) ++k;
    return k <= 15 ? k : -1;    // M >= 65536 corrupts the reference stream
}

extern "C" int drice_parse_cd_values(size_t n, const unsigned int *cd, drice_params *out)
{
    if (!out || (n > 0 && !cd)) return DRICE_E_PARAM;
    memset(out, 0, sizeof(*out));
    out->M = 8;                 // defaults, src/deltaRice.c:249-257
    out->L = -1;
    out->filter_len = 2;
    out->filter[0] = 1;
    out->filter[1] = -1;
    if (n >= 1) out->M = (int32_t)cd[0];
    if (n >= 2) out->L = (int32_t)cd[1];
    if (n >= 3) {
        out->filter_len = (int32_t)cd[2];
        if (out->filter_len <= 0 || (size_t)out->filter_len + 3 > n) return DRICE_E_PARAM;
        if (out->filter_len > DRICE_MAX_FILTER) return DRICE_E_UNSUPPORTED;
        for (int f = 0; f < out->filter_len; ++f) out->filter[f] = (int32_t)cd[3 + f];
        // decode divides by f[0] (src/deltaRice.c:100): the reference would trap (Appendix B9)
        if (out->filter[0] == 0) return DRICE_E_PARAM;
    }
    if (drice_log2_param(out->M) < 0) return DRICE_E_PARAM;
    if (out->L == 0 || out->L < -1) return DRICE_E_PARAM;
    return DRICE_OK;
}

extern "C" size_t drice_chunk_bound_bytes(size_t total, int64_t L)
{
    if (total == 0) return 4;
    size_t Lw = (L <= 0 || (uint64_t)L > total) ? total : (size_t)L;
    size_t W = (total + Lw - 1) / Lw;
    size_t tail = total - (W - 1) * Lw;
    return 4 * (1 + W + (W - 1) * wave_bound_words(Lw) + wave_bound_words(tail));
}

extern "C" size_t drice_batch_bound_bytes(const uint64_t *off, size_t nchunks, int64_t L)
{
    size_t b = 0;
    for (size_t c = 0; c < nchunks; ++c) b += drice_chunk_bound_bytes((size_t)(off[c + 1] - off[c]), L);
    return b;
}

// ======================================================================================
// context
// ======================================================================================
// One H2D and one D2H stream per device, shared by every handle of the process: bulk copies of
// the two directions then sit on two copy engines whatever the number of handles (with copy streams
// per handle, an encode batch and a decode batch running side by side serialised on the engines;
// measured on B200: 40 ms -> 30 ms for 1 GB each way).  Copies are only enqueued once they can run.
namespace {
std::mutex g_cs_mu;
cudaStream_t g_cs_in[64], g_cs_out[64];
bool copy_streams(int device, cudaStream_t *in, cudaStream_t *out)
{
    if (device < 0 || device >= 64) return false;
    std::lock_guard<std::mutex> lk(g_cs_mu);
    if (!g_cs_in[device]) {
        if (cudaStreamCreateWithFlags(&g_cs_in[device], cudaStreamNonBlocking) != cudaSuccess) { g_cs_in[device] = nullptr; return false; }
    }
    if (!g_cs_out[device]) {
        if (cudaStreamCreateWithFlags(&g_cs_out[device], cudaStreamNonBlocking) != cudaSuccess) { g_cs_out[device] = nullptr; return false; }
    }
    *in = g_cs_in[device];
    *out = g_cs_out[device];
    return true;
}
}  // namespace

extern "C" int drice_create(drice_ctx **out, int device)
{
    if (!out) return DRICE_E_PARAM;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, DRICE_E_CUDA, std::string("no usable CUDA device (this library has no CPU path): ") +
                                               cudaGetErrorString(e));
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= ndev) return fail(nullptr, DRICE_E_PARAM, "device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major < 10)
        return fail(nullptr, DRICE_E_CUDA, std::string("device ") + prop.name + " is not sm_100-class; kernels are built for sm_100a only");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    drice_ctx *ctx = new (std::nothrow) drice_ctx();
    if (!ctx) return fail(nullptr, DRICE_E_NOMEM, "out of memory");
    ctx->device = device;
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              copy_streams(device, &ctx->s_in, &ctx->s_out) &&
              cudaEventCreateWithFlags(&ctx->ev_tab, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_last, cudaEventDisableTiming) == cudaSuccess &&
              cudaMalloc((void **)&ctx->d_hint, 16) == cudaSuccess && cudaMemset(ctx->d_hint, 0, 16) == cudaSuccess &&
              cudaMallocHost((void **)&ctx->h_hint, 16) == cudaSuccess;
    if (ok) ctx->h_hint[0] = ctx->h_hint[1] = 0;
    for (Slot &s : ctx->slots)
        ok = ok && cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.ev_k, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        std::string m = std::string("stream/event creation failed: ") + cudaGetErrorString(cudaGetLastError());
        drice_destroy(ctx);
        return fail(nullptr, DRICE_E_CUDA, m);
    }
    *out = ctx;
    return DRICE_OK;
}

extern "C" void drice_destroy(drice_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (Slot &s : ctx->slots) {
        s.raw.release(); s.comp.release(); s.offs.release();
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_k) cudaEventDestroy(s.ev_k);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
        if (s.h_offs_pinned) cudaFreeHost(s.h_offs_pinned);
    }
    for (auto &t : ctx->timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    ctx->d_tab.release(); ctx->d_scratch.release(); ctx->d_offs.release(); ctx->d_filt.release(); ctx->d_scan.release();
    for (int i = 0; i < 2; ++i) {
        if (ctx->h_piece[i]) cudaFreeHost(ctx->h_piece[i]);
        if (ctx->ev_piece[i]) cudaEventDestroy(ctx->ev_piece[i]);
    }
    if (ctx->h_tab) cudaFreeHost(ctx->h_tab);
    if (ctx->d_hint) cudaFree(ctx->d_hint);
    if (ctx->h_hint) cudaFreeHost((void *)ctx->h_hint);
    if (ctx->h_sync) cudaFreeHost(ctx->h_sync);
    if (ctx->ev_tab) cudaEventDestroy(ctx->ev_tab);
    if (ctx->ev_last) cudaEventDestroy(ctx->ev_last);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *drice_last_error(const drice_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" int drice_device(const drice_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" int drice_set_filter(drice_ctx *ctx, const int32_t *filter, int filter_len)
{
    if (!ctx) return DRICE_E_PARAM;
    if (!filter) {                                       // back to the default delta filter
        ctx->filter_mode = 0;
        ctx->filter_len = 2;
        ctx->filter[0] = 1;
        ctx->filter[1] = -1;
        return DRICE_OK;
    }
    if (filter_len < 1) return fail(ctx, DRICE_E_PARAM, "empty pre-filter");
    if (filter_len > DRICE_MAX_FILTER) return fail(ctx, DRICE_E_UNSUPPORTED, "pre-filter longer than DRICE_MAX_FILTER taps");
    if (filter[0] == 0) return fail(ctx, DRICE_E_PARAM, "pre-filter tap 0 must not be 0 (decode divides by it)");
    ctx->filter_len = filter_len;
    for (int i = 0; i < filter_len; ++i) ctx->filter[i] = filter[i];
    // src/deltaRice.c:38-46 (checkIfDeltaFilter); [1] is the identity: nothing to run
    ctx->filter_mode = (filter_len == 2 && filter[0] == 1 && filter[1] == -1) ? 0 : (filter_len == 1 && filter[0] == 1) ? 1 : 2;
    return DRICE_OK;
}

extern "C" uint64_t drice_launch_count(const drice_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int drice_timing_enable(drice_ctx *ctx, int on)
{
    if (!ctx) return DRICE_E_PARAM;
    ctx->timing = on != 0;
    return DRICE_OK;
}

extern "C" int drice_timing_read(drice_ctx *ctx, double *ms, uint64_t *launches, int reset)
{
    if (!ctx) return DRICE_E_PARAM;
    DR_CUDA(ctx, cudaSetDevice(ctx->device));
    for (auto &t : ctx->timed) {
        DR_CUDA(ctx, cudaEventSynchronize(t.e1));
        float f = 0.f;
        DR_CUDA(ctx, cudaEventElapsedTime(&f, t.e0, t.e1));
        ctx->t_ms[t.kind] += f;
        ctx->t_n[t.kind] += 1;
        ctx->ev_pool.push_back(t.e0);
        ctx->ev_pool.push_back(t.e1);
    }
    ctx->timed.clear();
    for (int k = 0; k < DRICE_NUM_KERNELS; ++k) {
        if (ms) ms[k] = ctx->t_ms[k];
        if (launches) launches[k] = ctx->t_n[k];
        if (reset) { ctx->t_ms[k] = 0; ctx->t_n[k] = 0; }
    }
    return DRICE_OK;
}

extern "C" const char *drice_kernel_name(int kind)
{
    // kind 0 = the encode kernel of the batch: encode_tile_kernel / encode_multi_kernel
    static const char *names[DRICE_NUM_KERNELS] = {"encode_kernel", "locate_kernel", "parse_kernel"};
    return (kind >= 0 && kind < DRICE_NUM_KERNELS) ? names[kind] : nullptr;
}

extern "C" void *drice_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void drice_host_free(void *p) { if (p) cudaFreeHost(p); }

// ======================================================================================
// device entry points
// ======================================================================================
extern "C" int drice_encode_batch_dev_async(drice_ctx *ctx, const int16_t *d_raw, const uint64_t *off,
                                            size_t nchunks, int M, int64_t L, uint32_t *d_out,
                                            size_t out_cap_bytes, uint64_t *d_chunk_byte_off,
                                            uint32_t *d_status, void *stream)
{
    if (!ctx) return DRICE_E_PARAM;
    if (!off || !d_chunk_byte_off || !d_status) return fail(ctx, DRICE_E_PARAM, "null argument");
    const int k = drice_log2_param(M);
    if (k < 0) return fail(ctx, DRICE_E_PARAM, "RiceParameter must be a power of two in [1, 32768]");
    if ((reinterpret_cast<uintptr_t>(d_out) & 3) || (reinterpret_cast<uintptr_t>(d_raw) & 1))
        return fail(ctx, DRICE_E_PARAM, "misaligned buffer");
    DR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    Geometry g;
    int rc = build_geometry(ctx, off, nchunks, L, true, g);
    if (rc) return rc;
    if (nchunks == 0) {
        DR_CUDA(ctx, cudaMemsetAsync(d_status, 0, sizeof(uint32_t), st));
        DR_CUDA(ctx, cudaMemsetAsync(d_chunk_byte_off, 0, sizeof(uint64_t), st));
        return DRICE_OK;
    }
    if (!d_raw && off[nchunks] > 0) return fail(ctx, DRICE_E_PARAM, "null input");
    // scratch: [ticket u32][pad] | look-back status u64 per tile (at most one per wave): all zeroed
    const size_t zeroed = ((size_t)g.nwaves * 8 + 16 + 15) & ~(size_t)15;
    // + the segment tables of the long-wave encoder (few waves of more than 8192 samples; not zeroed)
    size_t long_bytes = g.nwaves <= 1024 ? encode_long_scratch_bytes(g.nwaves, g.max_wave) : 0;
    if (long_bytes > ((size_t)64 << 20)) long_bytes = 0;
    if (zeroed + long_bytes > ctx->d_scratch.cap) DR_CUDA(ctx, cudaDeviceSynchronize());
    DR_CUDA(ctx, ctx->d_scratch.reserve(zeroed + long_bytes));
    uint64_t *d_soff, *d_unused;
    uint32_t *d_woff;
    rc = upload_tables(ctx, off, nullptr, g.wave_off.data(), nchunks, st, &d_soff, &d_unused, &d_woff,
                       ctx->d_scratch.p, zeroed, d_status, true);
    if (rc) return rc;

    const int16_t *src = d_raw;
    if (ctx->filter_mode == 2) {
        // generic pre-filter (src/deltaRice.c:64-74): raw -> scratch, then coded with the delta off
        const size_t fb = (size_t)off[nchunks] * 2 + 64;
        if (fb > ctx->d_filt.cap) DR_CUDA(ctx, cudaDeviceSynchronize());
        DR_CUDA(ctx, ctx->d_filt.reserve(fb));
        FilterParams fp{};
        fp.chunk_sample_off = d_soff;
        fp.nchunks = (uint32_t)nchunks;
        fp.L = g.Lk;
        fp.flen = ctx->filter_len;
        for (int i = 0; i < ctx->filter_len; ++i) fp.f[i] = ctx->filter[i];
        uint64_t max_chunk = 0;
        for (size_t c = 0; c < nchunks; ++c) max_chunk = std::max<uint64_t>(max_chunk, off[c + 1] - off[c]);
        ctx->launches += (uint64_t)launch_prefilter(fp, d_raw, (int16_t *)ctx->d_filt.p, max_chunk, st);
        src = (const int16_t *)ctx->d_filt.p;
    }
    EncodeParams p{};
    p.raw = src;
    EncodeMode md{};
    md.delta = ctx->filter_mode == 0;
    // the hint is the largest record in WORDS of the previous batch; it only transfers to waves of the same shape
    if (ctx->hint_wave != g.max_wave || ctx->hint_k != k) ctx->h_hint[0] = 0;
    md.words_hint = ctx->h_hint[0];
    ctx->hint_wave = g.max_wave;
    ctx->hint_k = k;
    md.max_words = ctx->d_hint;
    md.long_scratch = long_bytes ? (char *)ctx->d_scratch.p + zeroed : nullptr;
    md.long_scratch_bytes = long_bytes;
    p.raw_samples = off[nchunks];
    p.out = d_out;
    p.out_cap_words = out_cap_bytes / 4;
    p.chunk_sample_off = d_soff;
    p.chunk_wave_off = d_woff;
    p.chunk_byte_off = d_chunk_byte_off;
    p.ticket = (uint32_t *)ctx->d_scratch.p;
    p.lookback = (uint64_t *)((char *)ctx->d_scratch.p + 16);
    p.status = d_status;
    p.nchunks = (uint32_t)nchunks;
    p.nwaves = g.nwaves;
    p.uniform_wpc = g.uniform_wpc;
    p.L = g.Lk;
    p.k = k;
    int nl;
    {
        TimedScope ts(ctx, DRICE_KERNEL_ENCODE, st);
        nl = launch_encode(p, md, g.max_wave, st);
    }
    if (nl < 0) return fail(ctx, DRICE_E_PARAM, "unsupported RiceParameter");
    ctx->launches += (uint64_t)nl;
    DR_CUDA(ctx, cudaGetLastError());
    DR_CUDA(ctx, cudaEventRecord(ctx->ev_last, st));
    ctx->last_stream = st;
    ctx->have_last = true;
    return DRICE_OK;
}

namespace {
int ensure_sync_buffers(drice_ctx *ctx, size_t nchunks)
{
    const size_t bytes = (nchunks + 1) * 8 + 8;
    DR_CUDA(ctx, ctx->d_offs.reserve(bytes));
    if (bytes > ctx->h_sync_cap) {
        if (ctx->h_sync) cudaFreeHost(ctx->h_sync);
        ctx->h_sync = nullptr;
        ctx->h_sync_cap = 0;
        DR_CUDA(ctx, cudaMallocHost((void **)&ctx->h_sync, bytes * 2));
        ctx->h_sync_cap = bytes * 2;
    }
    return DRICE_OK;
}
}  // namespace

extern "C" int drice_encode_batch_dev(drice_ctx *ctx, const int16_t *d_raw, const uint64_t *off,
                                      size_t nchunks, int M, int64_t L, uint32_t *d_out,
                                      size_t out_cap_bytes, uint64_t *chunk_byte_off, void *stream)
{
    if (!ctx) return DRICE_E_PARAM;
    if (!chunk_byte_off) return fail(ctx, DRICE_E_PARAM, "null argument");
    DR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc = ensure_sync_buffers(ctx, nchunks);
    if (rc) return rc;
    uint64_t *d_offs = (uint64_t *)ctx->d_offs.p;
    uint32_t *d_status = (uint32_t *)(d_offs + nchunks + 1);
    rc = drice_encode_batch_dev_async(ctx, d_raw, off, nchunks, M, L, d_out, out_cap_bytes, d_offs, d_status, st);
    if (rc) return rc;
    const size_t bytes = (nchunks + 1) * 8 + 8;
    if ((rc = small_copy(ctx, ctx->h_sync, d_offs, bytes, st))) return rc;
    DR_CUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t status = *(uint32_t *)(ctx->h_sync + nchunks + 1);
    memcpy(chunk_byte_off, ctx->h_sync, (nchunks + 1) * 8);
    return status_to_error(ctx, status);
}

extern "C" int drice_decode_batch_dev_async(drice_ctx *ctx, const uint32_t *d_comp, const uint64_t *boff,
                                            size_t nchunks, const uint64_t *off, int M, int64_t L,
                                            int16_t *d_out, uint32_t *d_status, void *stream)
{
    if (!ctx) return DRICE_E_PARAM;
    if (!boff || !off || !d_status) return fail(ctx, DRICE_E_PARAM, "null argument");
    const int k = drice_log2_param(M);
    if (k < 0) return fail(ctx, DRICE_E_PARAM, "RiceParameter must be a power of two in [1, 32768]");
    if ((reinterpret_cast<uintptr_t>(d_comp) & 3) || (reinterpret_cast<uintptr_t>(d_out) & 1))
        return fail(ctx, DRICE_E_PARAM, "misaligned buffer");
    DR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    Geometry g;
    int rc = build_geometry(ctx, off, nchunks, L, false, g);
    if (rc) return rc;
    if (nchunks == 0) {
        DR_CUDA(ctx, cudaMemsetAsync(d_status, 0, sizeof(uint32_t), st));
        return DRICE_OK;
    }
    std::vector<uint64_t> woff(nchunks + 1);
    for (size_t c = 0; c <= nchunks; ++c) {
        if ((boff[c] & 3) || (c && boff[c] < boff[c - 1] + 4))
            return fail(ctx, DRICE_E_PARAM, "chunk_byte_off must be 4-byte multiples, each chunk >= 4 bytes");
        woff[c] = boff[c] / 4;
    }
    if (!d_comp) return fail(ctx, DRICE_E_PARAM, "null input");
    const size_t nw = g.nwaves;
    // scratch: [parse ticket, 16 B][chain state of the long-wave parser] (zeroed) | wave tables
    const size_t long_bytes = parse_long_state_bytes(g.nwaves, g.max_wave);
    // [.. | 16 counters of the density sort] (zeroed) | wave tables | the sort's permutation
    const size_t zeroed = 16 + long_bytes + 64;
    // batches whose records are heavy on average (more than k + 5 bits per sample: escapes) are decoded with
    // waves of similar density sharing a warp (drice_decode.cu: wave_hist_kernel / wave_scatter_kernel)
    static const int sort_env = [] { const char *v = getenv("DRICE_DEC_SORT"); return v ? atoi(v) : 1; }();
    const bool dense = sort_env == 2 || (sort_env == 1 && (woff[nchunks] - woff[0]) * 32ull > (off[nchunks] - off[0]) * (uint64_t)(k + 5));
    const bool heavy_ok = dense || (woff[nchunks] - woff[0]) * 256ull > (off[nchunks] - off[0]) * (uint64_t)(8 * k + 28);   // (k + 3.5 bits: an escape in most groups of codes)
    const size_t scratch = zeroed + nw * (8 + 8 + 4) + 64 + 16 + (dense ? nw * 4 + 16 : 0);
    if (scratch > ctx->d_scratch.cap) DR_CUDA(ctx, cudaDeviceSynchronize());
    DR_CUDA(ctx, ctx->d_scratch.reserve(scratch));
    uint64_t *d_soff, *d_woff64;
    uint32_t *d_wave_off;
    rc = upload_tables(ctx, off, woff.data(), g.wave_off.data(), nchunks, st, &d_soff, &d_woff64, &d_wave_off,
                       ctx->d_scratch.p, zeroed, d_status);
    if (rc) return rc;
    uint64_t *wave_in = (uint64_t *)((char *)ctx->d_scratch.p + zeroed);
    uint64_t *wave_out = wave_in + nw;
    uint32_t *wave_n = (uint32_t *)(wave_out + nw);

    LocateParams lp{};
    lp.comp = d_comp;
    lp.comp_words = woff[nchunks];
    lp.chunk_word_off = d_woff64;
    lp.chunk_sample_off = d_soff;
    lp.chunk_wave_off = d_wave_off;
    lp.wave_in = wave_in;
    lp.wave_out = wave_out;
    lp.wave_n = wave_n;
    lp.status = d_status;
    lp.nchunks = (uint32_t)nchunks;
    lp.L = g.Lk;
    {
        uint64_t max_chunk_words = 0, max_chunk_waves = 0;
        for (size_t c = 0; c < nchunks; ++c) {
            max_chunk_words = std::max<uint64_t>(max_chunk_words, woff[c + 1] - woff[c]);
            max_chunk_waves = std::max<uint64_t>(max_chunk_waves, g.wave_off[c + 1] - g.wave_off[c]);
        }
        size_t scan_bytes = 0;
        const bool direct = g.Lk != 0 && locate_direct_applies(woff[nchunks] - woff[0], max_chunk_waves, nchunks);
        const bool scan = !direct && locate_scan_applies(g.Lk, k, max_chunk_words, nchunks, &scan_bytes);
        if (scan) {
            if (scan_bytes > ctx->d_scan.cap) DR_CUDA(ctx, cudaDeviceSynchronize());
            DR_CUDA(ctx, ctx->d_scan.reserve(scan_bytes));
        }
        TimedScope ts(ctx, DRICE_KERNEL_LOCATE, st);
        ctx->launches += (uint64_t)(direct ? launch_locate_direct(lp, k, st) : scan ? launch_locate_scan(lp, k, max_chunk_words, max_chunk_waves, ctx->d_scan.p, st) : launch_locate(lp, st));
    }

    ParseParams pp{};
    pp.comp = d_comp;
    pp.comp_words = woff[nchunks];
    pp.wave_in = wave_in;
    pp.wave_out = wave_out;
    pp.wave_n = wave_n;
    pp.out = d_out;
    pp.status = d_status;
    pp.ticket = (uint32_t *)ctx->d_scratch.p;
    pp.nwaves = g.nwaves;
    pp.max_n = g.max_wave;
    pp.k = k;
    pp.identity = ctx->filter_mode != 0;
    pp.long_state = long_bytes ? (unsigned long long *)((char *)ctx->d_scratch.p + 16) : nullptr;
    pp.sort_counters = (uint32_t *)((char *)ctx->d_scratch.p + 16 + long_bytes);
    pp.heavy_ok = heavy_ok ? 1 : 0;
    pp.sort_perm = dense ? (uint32_t *)(((uintptr_t)(wave_n + nw) + 15) & ~(uintptr_t)15) : nullptr;
    // widest store that every wave start allows
    int store_bytes = (int)(g.align_samples * 2);
    while (store_bytes > 2 && (reinterpret_cast<uintptr_t>(d_out) & (uintptr_t)(store_bytes - 1))) store_bytes >>= 1;
    if (store_bytes == 4) store_bytes = 2;
    int nl;
    {
        TimedScope ts(ctx, DRICE_KERNEL_PARSE, st);
        nl = launch_parse(pp, store_bytes, st);
    }
    if (nl < 0) return fail(ctx, DRICE_E_PARAM, "unsupported RiceParameter");
    ctx->launches += (uint64_t)nl;
    if (ctx->filter_mode == 2) {
        // inverse of the generic pre-filter (src/deltaRice.c:91-102), in place on the output
        FilterParams fp{};
        fp.chunk_sample_off = d_soff;
        fp.nchunks = (uint32_t)nchunks;
        fp.L = g.Lk;
        fp.flen = ctx->filter_len;
        for (int i = 0; i < ctx->filter_len; ++i) fp.f[i] = ctx->filter[i];
        uint64_t max_waves = 0;
        for (size_t c = 0; c < nchunks; ++c) max_waves = std::max<uint64_t>(max_waves, g.wave_off[c + 1] - g.wave_off[c]);
        ctx->launches += (uint64_t)launch_postfilter(fp, d_out, max_waves, st);
    }
    DR_CUDA(ctx, cudaGetLastError());
    DR_CUDA(ctx, cudaEventRecord(ctx->ev_last, st));
    ctx->last_stream = st;
    ctx->have_last = true;
    return DRICE_OK;
}

extern "C" int drice_decode_batch_dev(drice_ctx *ctx, const uint32_t *d_comp, const uint64_t *boff,
                                      size_t nchunks, const uint64_t *off, int M, int64_t L,
                                      int16_t *d_out, void *stream)
{
    if (!ctx) return DRICE_E_PARAM;
    DR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc = ensure_sync_buffers(ctx, 0);
    if (rc) return rc;
    uint32_t *d_status = (uint32_t *)ctx->d_offs.p;
    rc = drice_decode_batch_dev_async(ctx, d_comp, boff, nchunks, off, M, L, d_out, d_status, st);
    if (rc) return rc;
    if ((rc = small_copy(ctx, ctx->h_sync, d_status, 4, st))) return rc;
    DR_CUDA(ctx, cudaStreamSynchronize(st));
    return status_to_error(ctx, *(uint32_t *)ctx->h_sync);
}

// ======================================================================================
// chunk scheduler: host-pointer entry points
// ======================================================================================
namespace {

// ---- pageable caller buffers ---------------------------------------------------------
// libhdf5 hands the filter malloc'ed (pageable) memory.  cudaMemcpyAsync on pageable memory is a
// synchronous staged copy at ~12 GB/s; PCIe 5 x16 moves 55.  Large pageable transfers are therefore
// staged HERE: 4 MB pieces through two pinned buffers, the memcpy of a piece split over a small
// pool of host threads, piece i+1 being copied while piece i is on the bus.
constexpr size_t kPieceBytes = 4u << 20;
constexpr size_t kStagedMin = 8u << 20;      // smaller transfers: leave it to the driver

class CopyPool {
public:
    static CopyPool &get()
    {
        static CopyPool pool;
        return pool;
    }
    // memcpy split in kParts; returns when all of it is done
    void copy(char *dst, const char *src, size_t n)
    {
        if (!ok_ || n < (1u << 20)) {
            memcpy(dst, src, n);
            return;
        }
        const size_t part = (n / kParts + 63) & ~(size_t)63;
        int mine = 0;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (int i = 1; i < kParts; ++i) {
                const size_t lo = part * i;
                if (lo >= n) break;
                q_.push_back({dst + lo, src + lo, std::min(part, n - lo)});
                ++mine;
            }
            pending_ += mine;
        }
        cv_.notify_all();
        memcpy(dst, src, std::min(part, n));
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    static constexpr int kParts = 4;
    struct Job { char *dst; const char *src; size_t n; };
    CopyPool()
    {
        try {
            for (int i = 0; i < kParts - 1; ++i) th_.emplace_back([this] { run(); });
            ok_ = true;
        } catch (...) {
            ok_ = false;
        }
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void run()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                j = q_.front();
                q_.pop_front();
            }
            memcpy(j.dst, j.src, j.n);
            {
                std::lock_guard<std::mutex> lk(m_);
                --pending_;
            }
            done_.notify_all();
        }
    }
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::deque<Job> q_;
    std::vector<std::thread> th_;
    int pending_ = 0;
    bool stop_ = false, ok_ = false;
};

bool is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

int ensure_pieces(drice_ctx *ctx)
{
    for (int i = 0; i < 2; ++i) {
        if (!ctx->h_piece[i]) DR_CUDA(ctx, cudaMallocHost((void **)&ctx->h_piece[i], kPieceBytes));
        if (!ctx->ev_piece[i]) DR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_piece[i], cudaEventDisableTiming));
    }
    return DRICE_OK;
}

// host -> device on `st`; the source may be reused when the call returns iff it was staged
// (pageable); pinned sources follow the usual asynchronous rules
int h2d(drice_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st)
{
    if (bytes < kStagedMin || !is_pageable(h_src)) {
        DR_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return DRICE_OK;
    }
    int rc = ensure_pieces(ctx);
    if (rc) return rc;
    size_t off = 0;
    for (int i = 0; off < bytes; ++i) {
        const int sl = i & 1;
        const size_t n = std::min(kPieceBytes, bytes - off);
        if (i >= 2) DR_CUDA(ctx, cudaEventSynchronize(ctx->ev_piece[sl]));
        CopyPool::get().copy(ctx->h_piece[sl], (const char *)h_src + off, n);
        DR_CUDA(ctx, cudaMemcpyAsync((char *)d_dst + off, ctx->h_piece[sl], n, cudaMemcpyHostToDevice, st));
        DR_CUDA(ctx, cudaEventRecord(ctx->ev_piece[sl], st));
        off += n;
    }
    // the pieces are reused by the next staged transfer: drain
    DR_CUDA(ctx, cudaEventSynchronize(ctx->ev_piece[0]));
    DR_CUDA(ctx, cudaEventSynchronize(ctx->ev_piece[1]));
    return DRICE_OK;
}

// device -> host on `st`.  Returns true in *done when the data is already in h_dst (staged path:
// the call waited for it); otherwise the copy is in flight on `st` as usual
int d2h(drice_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, cudaStream_t st, bool *done)
{
    *done = false;
    if (bytes < kStagedMin || !is_pageable(h_dst)) {
        DR_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        return DRICE_OK;
    }
    int rc = ensure_pieces(ctx);
    if (rc) return rc;
    const size_t npieces = (bytes + kPieceBytes - 1) / kPieceBytes;
    auto issue = [&](size_t i) -> cudaError_t {
        const size_t off = i * kPieceBytes, n = std::min(kPieceBytes, bytes - off);
        cudaError_t e = cudaMemcpyAsync(ctx->h_piece[i & 1], (const char *)d_src + off, n, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(ctx->ev_piece[i & 1], st);
    };
    DR_CUDA(ctx, issue(0));
    for (size_t i = 0; i < npieces; ++i) {
        DR_CUDA(ctx, cudaEventSynchronize(ctx->ev_piece[i & 1]));
        if (i + 1 < npieces) DR_CUDA(ctx, issue(i + 1));          // the other piece is free: its copy-out was i-1
        const size_t off = i * kPieceBytes, n = std::min(kPieceBytes, bytes - off);
        CopyPool::get().copy((char *)h_dst + off, ctx->h_piece[i & 1], n);
    }
    *done = true;
    return DRICE_OK;
}

size_t subbatch_bytes()
{
    static size_t v = 0;
    if (!v) {
        const char *e = getenv("DRICE_SUBBATCH_MB");
        long mb = e ? atol(e) : 64;
        if (mb < 1) mb = 1;
        v = (size_t)mb << 20;
    }
    return v;
}

int slot_offs_reserve(drice_ctx *ctx, Slot &s, size_t nchunks)
{
    const size_t bytes = (nchunks + 1) * 8 + 8;
    DR_CUDA(ctx, s.offs.reserve(bytes));
    if (bytes > s.h_offs_cap) {
        if (s.h_offs_pinned) cudaFreeHost(s.h_offs_pinned);
        s.h_offs_pinned = nullptr;
        s.h_offs_cap = 0;
        DR_CUDA(ctx, cudaMallocHost((void **)&s.h_offs_pinned, bytes * 2));
        s.h_offs_cap = bytes * 2;
    }
    return DRICE_OK;
}

// cuts [0,nchunks) into runs of whole chunks of about `target` raw bytes
std::vector<size_t> cut_subbatches(const uint64_t *off, size_t nchunks, size_t target)
{
    std::vector<size_t> cuts{0};
    size_t c = 0;
    while (c < nchunks) {
        size_t e = c + 1;
        while (e < nchunks && (off[e + 1] - off[c]) * 2 <= target) ++e;
        cuts.push_back(e);
        c = e;
    }
    return cuts;
}

}  // namespace

extern "C" int drice_encode_batch_host(drice_ctx *ctx, const int16_t *h_raw, const uint64_t *off,
                                       size_t nchunks, int M, int64_t L, void *h_out,
                                       size_t out_cap_bytes, uint64_t *chunk_byte_off)
{
    if (!ctx) return DRICE_E_PARAM;
    if (!off || !chunk_byte_off || (!h_out && out_cap_bytes)) return fail(ctx, DRICE_E_PARAM, "null argument");
    if (drice_log2_param(M) < 0) return fail(ctx, DRICE_E_PARAM, "RiceParameter must be a power of two in [1, 32768]");
    DR_CUDA(ctx, cudaSetDevice(ctx->device));
    chunk_byte_off[0] = 0;
    if (nchunks == 0) return DRICE_OK;
    const std::vector<size_t> cuts = cut_subbatches(off, nchunks, subbatch_bytes());
    const size_t nsub = cuts.size() - 1;
    constexpr int NS = 3;
    uint64_t out_pos = 0;      // bytes written to h_out so far (known once a sub-batch's kernel is done)
    int rc = DRICE_OK;

    // software pipeline: stage i = [H2D + kernels + offsets D2H] ; stage i-1 = [wait kernels,
    // issue stream D2H at its final position]
    auto finish = [&](size_t i) -> int {
        Slot &s = ctx->slots[i % NS];
        DR_CUDA(ctx, cudaEventSynchronize(s.ev_k));
        const size_t n = s.c1 - s.c0;
        const uint32_t status = *(uint32_t *)(s.h_offs_pinned + n + 1);
        int r = status_to_error(ctx, status);
        if (r) return r;
        const uint64_t bytes = s.h_offs_pinned[n];
        if (out_pos + bytes > out_cap_bytes) return fail(ctx, DRICE_E_CAPACITY, "output buffer too small for the compressed batch");
        for (size_t c = 0; c <= n; ++c) chunk_byte_off[s.c0 + c] = out_pos + s.h_offs_pinned[c];
        bool landed = false;
        { int r2 = d2h(ctx, (char *)h_out + out_pos, s.comp.p, bytes, ctx->s_out, &landed); if (r2) return r2; }
        DR_CUDA(ctx, cudaEventRecord(s.ev_out, ctx->s_out));
        out_pos += bytes;
        return DRICE_OK;
    };

    for (size_t i = 0; i < nsub + 1 && rc == DRICE_OK; ++i) {
        if (i < nsub) {
            Slot &s = ctx->slots[i % NS];
            if (s.busy) {                                  // slot's previous D2H must have drained
                DR_CUDA(ctx, cudaEventSynchronize(s.ev_out));
                s.busy = false;
            }
            s.c0 = cuts[i];
            s.c1 = cuts[i + 1];
            const size_t n = s.c1 - s.c0;
            const uint64_t s0 = off[s.c0], s1 = off[s.c1];
            s.h_offs.assign(off + s.c0, off + s.c1 + 1);
            for (uint64_t &v : s.h_offs) v -= s0;
            const size_t bound = drice_batch_bound_bytes(s.h_offs.data(), n, L);
            if (s.raw.cap < (s1 - s0) * 2 + 64 || s.comp.cap < bound) DR_CUDA(ctx, cudaDeviceSynchronize());
            DR_CUDA(ctx, s.raw.reserve((s1 - s0) * 2 + 64));
            DR_CUDA(ctx, s.comp.reserve(bound));
            if ((rc = slot_offs_reserve(ctx, s, n))) break;
            if ((rc = h2d(ctx, s.raw.p, h_raw + s0, (s1 - s0) * 2, ctx->s_in))) break;
            DR_CUDA(ctx, cudaEventRecord(s.ev_in, ctx->s_in));
            DR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s.ev_in, 0));
            uint64_t *d_offs = (uint64_t *)s.offs.p;
            uint32_t *d_status = (uint32_t *)(d_offs + n + 1);
            rc = drice_encode_batch_dev_async(ctx, (const int16_t *)s.raw.p, s.h_offs.data(), n, M, L,
                                              (uint32_t *)s.comp.p, s.comp.cap, d_offs, d_status, ctx->stream);
            if (rc) break;
            if ((rc = small_copy(ctx, s.h_offs_pinned, d_offs, (n + 1) * 8 + 8, ctx->stream))) break;
            DR_CUDA(ctx, cudaEventRecord(s.ev_k, ctx->stream));
            s.busy = true;
        }
        if (i >= 1) rc = finish(i - 1);
    }
    // drain (own events: the copy streams are shared with other handles)
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    for (Slot &s : ctx->slots) {
        if (s.busy) {
            cudaError_t e2 = cudaEventSynchronize(s.ev_out);
            if (e == cudaSuccess) e = e2;
        }
        s.busy = false;
    }
    if (rc == DRICE_OK && e != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamSynchronize");
    return rc;
}

extern "C" int drice_decode_batch_host(drice_ctx *ctx, const void *h_comp, const uint64_t *boff,
                                       size_t nchunks, const uint64_t *off, int M, int64_t L,
                                       int16_t *h_out)
{
    if (!ctx) return DRICE_E_PARAM;
    if (!boff || !off) return fail(ctx, DRICE_E_PARAM, "null argument");
    if (drice_log2_param(M) < 0) return fail(ctx, DRICE_E_PARAM, "RiceParameter must be a power of two in [1, 32768]");
    DR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (nchunks == 0) return DRICE_OK;
    const std::vector<size_t> cuts = cut_subbatches(off, nchunks, subbatch_bytes());
    const size_t nsub = cuts.size() - 1;
    constexpr int NS = 3;
    int rc = DRICE_OK;

    // software pipeline: stage i = [stream H2D + kernels]; stage i-1 = [wait kernels, issue the raw
    // D2H].  A copy is only handed to a copy engine once it can run: an engine is a FIFO shared by
    // every stream of the process, and a copy parked at its head on a stream dependency would block
    // the transfers of other handles behind it (encode and decode batches side by side).
    auto finish = [&](size_t i) -> int {
        Slot &s = ctx->slots[i % NS];
        DR_CUDA(ctx, cudaEventSynchronize(s.ev_k));
        int r = status_to_error(ctx, *(uint32_t *)s.h_offs_pinned);
        if (r) return r;
        const uint64_t s0 = off[s.c0], s1 = off[s.c1];
        bool landed = false;
        { int r2 = d2h(ctx, h_out + s0, s.raw.p, (s1 - s0) * 2, ctx->s_out, &landed); if (r2) return r2; }
        DR_CUDA(ctx, cudaEventRecord(s.ev_out, ctx->s_out));
        return DRICE_OK;
    };

    for (size_t i = 0; i < nsub + 1 && rc == DRICE_OK; ++i) {
        if (i < nsub) {
            Slot &s = ctx->slots[i % NS];
            if (s.busy) {                                  // slot's previous D2H must have drained
                DR_CUDA(ctx, cudaEventSynchronize(s.ev_out));
                s.busy = false;
            }
            s.c0 = cuts[i];
            s.c1 = cuts[i + 1];
            const size_t n = s.c1 - s.c0;
            const uint64_t s0 = off[s.c0], s1 = off[s.c1];
            const uint64_t b0 = boff[s.c0], b1 = boff[s.c1];
            if (b1 < b0 || (b0 & 3) || (b1 & 3)) { rc = fail(ctx, DRICE_E_PARAM, "chunk_byte_off must be non-decreasing 4-byte multiples"); break; }
            s.h_offs.resize(2 * (n + 1));
            for (size_t c = 0; c <= n; ++c) {
                s.h_offs[c] = off[s.c0 + c] - s0;
                s.h_offs[n + 1 + c] = boff[s.c0 + c] - b0;
            }
            if (s.raw.cap < (s1 - s0) * 2 + 64 || s.comp.cap < (b1 - b0) + 64) DR_CUDA(ctx, cudaDeviceSynchronize());
            DR_CUDA(ctx, s.raw.reserve((s1 - s0) * 2 + 64));
            DR_CUDA(ctx, s.comp.reserve((b1 - b0) + 64));
            if ((rc = slot_offs_reserve(ctx, s, 1))) break;
            if ((rc = h2d(ctx, s.comp.p, (const char *)h_comp + b0, b1 - b0, ctx->s_in))) break;
            DR_CUDA(ctx, cudaEventRecord(s.ev_in, ctx->s_in));
            DR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s.ev_in, 0));
            uint32_t *d_status = (uint32_t *)s.offs.p;
            rc = drice_decode_batch_dev_async(ctx, (const uint32_t *)s.comp.p, s.h_offs.data() + n + 1, n,
                                              s.h_offs.data(), M, L, (int16_t *)s.raw.p, d_status, ctx->stream);
            if (rc) break;
            if ((rc = small_copy(ctx, s.h_offs_pinned, d_status, 4, ctx->stream))) break;
            DR_CUDA(ctx, cudaEventRecord(s.ev_k, ctx->stream));
            s.busy = true;
        }
        if (i >= 1) rc = finish(i - 1);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    for (Slot &s : ctx->slots) {
        if (s.busy) {
            cudaError_t e2 = cudaEventSynchronize(s.ev_out);
            if (e == cudaSuccess) e = e2;
        }
        s.busy = false;
    }
    if (rc == DRICE_OK && e != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamSynchronize");
    return rc;
}

extern "C" int drice_peek_chunk_samples(const void *h_comp, const uint64_t *boff, size_t nchunks,
                                        uint64_t *chunk_samples)
{
    if (!h_comp || !boff || !chunk_samples) return DRICE_E_PARAM;
    for (size_t c = 0; c < nchunks; ++c) {
        if (boff[c + 1] < boff[c] + 4) return DRICE_E_STREAM;
        uint32_t t;
        memcpy(&t, (const char *)h_comp + boff[c], 4);
        chunk_samples[c] = t;
    }
    return DRICE_OK;
}
