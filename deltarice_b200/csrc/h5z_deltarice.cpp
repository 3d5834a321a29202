// h5z_deltarice.cpp — the HDF5 boundary: filter class 32025, the filter callback, the
// registration helper and the dynamic-plugin entry points.
//
// Replaces (reference, paths relative to /root/reference):
//   H5Z_DELTARICE class struct         src/deltaRice.c:19-28
//   H5Z_filter_deltarice               src/deltaRice.c:468-490
//   deltarice_register_h5filter        src/deltaRice.c:494-501
//   H5PLget_plugin_type / _info        src/deltaRice_h5plugin.c:4-5
//
// libhdf5 calls the filter once per chunk, synchronously: a batch of one chunk goes through
// the chunk scheduler (drice_*_batch_host).  There is no CPU codec behind this file: when no
// B200-class device is usable the callback fails (returns 0) and says why on stderr.
//
// Deliberate differences from the reference:
//   B1  H5PLget_plugin_info returns the class pointer (reference returns (void*)32025)
//   B2  failure returns 0 (reference returns (size_t)-1) and leaves *buf untouched
//   B5/B9/B10  invalid M, WaveformLength 0, odd byte counts, empty filters are rejected
//   generic pre-filters (cd_nelmts >= 3): up to 16 taps, first tap != 0 (the reference divides by it)
#include "../../include/deltaRice.h"
#include "../../include/deltarice_b200.h"
#ifdef DRICE_USE_SYSTEM_HDF5
#include "H5PLextern.h"
#else
#include "../../include/hdf5_abi/H5PLextern.h"
#endif

#include <dlfcn.h>
#include <link.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <sys/mman.h>
#include <thread>
#include <vector>

namespace {
std::mutex g_mu;
drice_ctx *g_ctx = nullptr;

drice_ctx *global_ctx()
{
    if (!g_ctx) {
        int dev = -1;
        if (const char *e = getenv("DRICE_DEVICE")) dev = atoi(e);
        if (drice_create(&g_ctx, dev) != DRICE_OK) {
            fprintf(stderr, "deltarice_b200: %s\n", drice_last_error(nullptr));
            g_ctx = nullptr;
        }
    }
    return g_ctx;
}
}  // namespace

extern "C" {

H5Z_class_t H5Z_DELTARICE[1] = {{
    H5Z_CLASS_T_VERS,                       /* H5Z_class_t version        */
    (H5Z_filter_t)H5Z_FILTER_DELTARICE,     /* filter id 32025            */
    1,                                      /* encoder present            */
    1,                                      /* decoder present            */
    "deltarice",                            /* name (as the reference's)  */
    NULL,                                   /* can_apply                  */
    NULL,                                   /* set_local                  */
    (H5Z_func_t)H5Z_filter_deltarice,
}};

namespace {
// The buffer handed back to libhdf5 must come from malloc (libhdf5 frees it).  For a multi-MB chunk
// glibc maps fresh pages, and faulting them in one by one inside the device-to-host copy costs more
// than the codec (28 MB: ~10 ms against ~1 ms of GPU work).  Ask for huge pages and populate the
// pages from a few threads WHILE the stream goes to the device and the kernels run:
// MADV_POPULATE_WRITE faults pages in without touching their contents, so it cannot race with the
// copy that fills the buffer afterwards (kernels older than 5.14: touch the pages up front instead).
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
struct Prefault {
    std::vector<std::thread> th;
    void *p = nullptr;
    void start(size_t bytes)
    {
        p = malloc(bytes ? bytes : 1);
        if (!p || bytes < (4u << 20)) return;
        const uintptr_t a = ((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
        const uintptr_t e = ((uintptr_t)p + bytes) & ~(uintptr_t)((2u << 20) - 1);
        if (e > a) madvise((void *)a, e - a, MADV_HUGEPAGE);
        // probe: is MADV_POPULATE_WRITE there?
        const uintptr_t pg = ((uintptr_t)p + 4095) & ~(uintptr_t)4095;
        const bool populate = madvise((void *)pg, 4096, MADV_POPULATE_WRITE) == 0;
        const unsigned nt = 4;
        void *base = p;
        try {                                            // (no exception may cross the C boundary)
            for (unsigned t = 0; t < nt; ++t)
                th.emplace_back([=] {
                    const size_t lo = bytes / nt * t, hi = (t + 1 == nt) ? bytes : bytes / nt * (t + 1);
                    if (populate) {
                        const uintptr_t l = ((uintptr_t)base + lo + 4095) & ~(uintptr_t)4095;
                        const uintptr_t h = ((uintptr_t)base + hi) & ~(uintptr_t)4095;
                        if (h > l) madvise((void *)l, h - l, MADV_POPULATE_WRITE);
                    } else {
                        volatile char *q = (volatile char *)base;
                        for (size_t i = lo; i < hi; i += 4096) q[i] = 0;
                    }
                });
        } catch (...) {
        }
        if (!populate) join();                           // plain touches must be done before the data arrives
    }
    void join()
    {
        for (auto &t : th) t.join();
        th.clear();
    }
    ~Prefault() { join(); }
};

// memcpy into freshly malloc'ed memory, page faults spread over a few threads for multi-MB streams
void copy_out(void *dst, const void *src, size_t bytes)
{
    if (bytes < (4u << 20)) {
        memcpy(dst, src, bytes);
        return;
    }
    const unsigned nt = 4;
    std::vector<std::thread> th;
    unsigned started = 0;
    try {                                                // (no exception may cross the C boundary)
        for (unsigned t = 0; t < nt; ++t) {
            th.emplace_back([=] {
                const size_t lo = bytes / nt * t, hi = (t + 1 == nt) ? bytes : bytes / nt * (t + 1);
                memcpy((char *)dst + lo, (const char *)src + lo, hi - lo);
            });
            ++started;
        }
    } catch (...) {
    }
    for (auto &t : th) t.join();
    if (started < nt) {                                  // the parts no thread took
        const size_t lo = bytes / nt * started;
        memcpy((char *)dst + lo, (const char *)src + lo, bytes - lo);
    }
}

void  *g_stage = nullptr;        // pinned landing zone of the encoder's output (guarded by g_mu)
size_t g_stage_cap = 0;
}  // namespace

size_t H5Z_filter_deltarice(unsigned flags, size_t cd_nelmts, const unsigned cd_values[],
                            size_t nbytes, size_t *buf_size, void **buf)
{
    if (!buf || !*buf || !buf_size) return 0;
    drice_params prm;
    const int prc = drice_parse_cd_values(cd_nelmts, cd_values, &prm);
    if (prc != DRICE_OK) {
        fprintf(stderr, prc == DRICE_E_UNSUPPORTED
                            ? "deltarice_b200: pre-filters of more than 16 taps are not supported\n"
                            : "deltarice_b200: invalid compression_opts (RiceParameter must be 2^k <= 32768, WaveformLength >= 1 or -1, "
                              "filter length >= 1, first tap != 0)\n");
        return 0;
    }
    std::lock_guard<std::mutex> lock(g_mu);
    drice_ctx *ctx = global_ctx();
    if (!ctx) return 0;
    // the pre-filter of THIS call (cd_values[2..], src/deltaRice.c:277-290); the handle is shared by
    // every dataset of the process and calls are serialised by g_mu
    if (drice_set_filter(ctx, prm.filter, prm.filter_len) != DRICE_OK) {
        fprintf(stderr, "deltarice_b200: %s\n", drice_last_error(ctx));
        return 0;
    }

    if (flags & H5Z_FLAG_REVERSE) {
        if (nbytes < 4 || (nbytes & 3)) {
            fprintf(stderr, "deltarice_b200: compressed chunk size %zu is not a positive multiple of 4\n", nbytes);
            return 0;
        }
        uint32_t total;
        memcpy(&total, *buf, 4);
        if (total > 0x7fffffffu) return 0;
        const size_t out_bytes = (size_t)total * 2;
        Prefault pf;
        pf.start(out_bytes);
        void *out = pf.p;
        if (!out) return 0;
        const uint64_t boff[2] = {0, nbytes};
        const uint64_t soff[2] = {0, total};
        const int rc = drice_decode_batch_host(ctx, *buf, boff, 1, soff, prm.M, prm.L, (int16_t *)out);
        pf.join();
        if (rc != DRICE_OK) {
            fprintf(stderr, "deltarice_b200: de-compression failed: %s\n", drice_last_error(ctx));
            free(out);
            return 0;
        }
        free(*buf);
        *buf = out;
        *buf_size = out_bytes;
        return out_bytes;
    }

    if (nbytes & 1) {
        fprintf(stderr, "deltarice_b200: chunk of %zu bytes is not a whole number of int16 samples\n", nbytes);
        return 0;
    }
    const size_t total = nbytes / 2;
    if (total > 0x7fffffffull) return 0;
    // the stream's size is only known afterwards: encode into a pinned scratch that lives with the
    // plugin (sized for the worst case), then hand libhdf5 a malloc'ed buffer of exactly the size
    const size_t bound = drice_chunk_bound_bytes(total, prm.L);
    if (bound > g_stage_cap) {
        if (g_stage) drice_host_free(g_stage);
        g_stage_cap = 0;
        g_stage = drice_host_alloc(bound + bound / 4);
        if (!g_stage) {
            fprintf(stderr, "deltarice_b200: out of pinned memory (%zu bytes)\n", bound);
            return 0;
        }
        g_stage_cap = bound + bound / 4;
    }
    const uint64_t soff[2] = {0, total};
    uint64_t boff[2] = {0, 0};
    const int rc = drice_encode_batch_host(ctx, (const int16_t *)*buf, soff, 1, prm.M, prm.L, g_stage, bound, boff);
    if (rc != DRICE_OK) {
        fprintf(stderr, "deltarice_b200: compression failed: %s\n", drice_last_error(ctx));
        return 0;
    }
    const size_t used = (size_t)boff[1];
    void *out = malloc(used ? used : 1);
    if (!out) return 0;
    copy_out(out, g_stage, used);
    free(*buf);
    *buf = out;
    *buf_size = used;
    return used;
}

int deltarice_register_h5filter(void)
{
    typedef herr_t (*H5Zregister_t)(const void *);
    // libhdf5 is resolved from the running process (no link-time dependency).  The global scope first
    // (C programs linked against libhdf5, LD_PRELOAD); then every loaded object whose name contains
    // "libhdf5": Python extension modules are loaded RTLD_LOCAL and pull in a private, often hashed
    // libhdf5-*.so whose symbols RTLD_DEFAULT does not see.
    H5Zregister_t reg = (H5Zregister_t)dlsym(RTLD_DEFAULT, "H5Zregister");
    if (!reg) {
        struct Find { H5Zregister_t fn; } find{nullptr};
        dl_iterate_phdr([](struct dl_phdr_info *info, size_t, void *data) -> int {
            Find *f = (Find *)data;
            const char *name = info->dlpi_name;
            if (!name || !strstr(name, "libhdf5") || strstr(name, "libhdf5_hl")) return 0;
            if (void *h = dlopen(name, RTLD_NOLOAD | RTLD_LAZY)) {
                f->fn = (H5Zregister_t)dlsym(h, "H5Zregister");
                dlclose(h);                              // (RTLD_NOLOAD took a reference)
            }
            return f->fn != nullptr;
        }, &find);
        reg = find.fn;
    }
    if (!reg) {
        fprintf(stderr, "deltarice_register_h5filter: no libhdf5 in this process (H5Zregister not found)\n");
        return -1;
    }
    const int retval = reg(H5Z_DELTARICE);
    if (retval < 0) fprintf(stderr, "deltarice_register_h5filter: can't register deltarice filter\n");
    return retval;
}

H5PL_type_t H5PLget_plugin_type(void) { return H5PL_TYPE_FILTER; }
const void *H5PLget_plugin_info(void) { return H5Z_DELTARICE; }

}  // extern "C"
