/* deltarice_b200.h — thin C-ABI of the B200-native Delta-Rice codec.
 *
 * This is the boundary a host-language binding (cgo / JNI / ctypes / the HDF5 filter in
 * include/deltaRice.h) binds to.  Plain pointers and sizes only; no torch / C++ types.
 * Everything here runs on hand-written sm_100a CUDA kernels; there is NO CPU fallback:
 * every entry point returns an error when no CUDA device is usable.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   drice_encode_batch_*   <- writeWholeCompressedByteString  src/deltaRice.c:383-465
 *                             (+ perWaveCompression :365-381, encodeWaveform :49-63,
 *                              compressWithRiceCoding :191-244), called once per chunk
 *   drice_decode_batch_*   <- readWholeCompressedByteString   src/deltaRice.c:301-358
 *                             (+ perWaveDecompression :293-297, decompressWithRiceCoding
 *                              :138-189, decodeWaveform :78-90)
 *   drice_set_filter       <- the `filter` argument of encodeWaveform / decodeWaveform
 *                             src/deltaRice.c:49-104 (generic branches :64-74, :91-102)
 *   drice_parse_cd_values  <- parseCD_VALUES                  src/deltaRice.c:248-291
 *   drice_log2_param       <- determinePowerOf2               src/deltaRice.c:114-136
 * The reference handles ONE chunk per call (libhdf5 calls the filter per chunk); the batch
 * entry points take N chunks per call so one launch covers many HDF5 chunks (the "chunk
 * scheduler" of BASELINE.json north_star (3)).  A batch of one is the H5Z path.
 *
 * Stream format (bit-exact with the reference):
 *   chunk  := u32 total_samples, record[0..W-1]      W = ceil(total/L)
 *   record := u32 nwords, u32 word[nwords]           MSB-first Rice codes of one wave
 *
 * Error convention: functions returning int give 0 on success, a negative DRICE_E_* code
 * otherwise; drice_last_error(ctx) has the text.
 */
#ifndef DELTARICE_B200_H
#define DELTARICE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRICE_ABI_VERSION 2

#if defined(__GNUC__)
#define DRICE_API __attribute__((visibility("default")))
#else
#define DRICE_API
#endif

#define DRICE_OK            0
#define DRICE_E_PARAM      -1   /* bad M / L / cd_values / sizes                        */
#define DRICE_E_CUDA       -2   /* CUDA runtime error or no usable device               */
#define DRICE_E_CAPACITY   -3   /* output buffer too small                              */
#define DRICE_E_STREAM     -4   /* malformed compressed stream                          */
#define DRICE_E_NOMEM      -5
#define DRICE_E_UNSUPPORTED -6  /* pre-filter longer than DRICE_MAX_FILTER taps           */

#define DRICE_MAX_FILTER   16   /* taps of a generic pre-filter (cd_values[3..])        */

typedef struct drice_ctx drice_ctx;

/* Parsed compression_opts / cd_values (reference parseCD_VALUES, src/deltaRice.c:248-291). */
typedef struct drice_params {
    int32_t M;          /* Rice parameter, 2^k, 1 <= M <= 32768                        */
    int32_t L;          /* WaveformLength in samples; -1 = whole chunk is one wave      */
    int32_t filter_len; /* 2 for the default delta filter [1,-1]                        */
    int32_t filter[DRICE_MAX_FILTER];
} drice_params;

DRICE_API int drice_abi_version(void);

/* k = log2(M) for M = 2^k, 0 <= k <= 15; -1 otherwise (src/deltaRice.c:114-136). */
DRICE_API int drice_log2_param(int M);

/* cd_values -> params.  n = 0: M=8, L=-1;  n = 1: (M);  n = 2: (M, L);  n >= 3: (M, L,
 * filter_len, f0, f1, ...).  Returns DRICE_OK, DRICE_E_PARAM (bad M, L == 0, L < -1,
 * filter_len <= 0, f0 == 0) or DRICE_E_UNSUPPORTED (more than DRICE_MAX_FILTER taps). */
DRICE_API int drice_parse_cd_values(size_t cd_nelmts, const unsigned int *cd_values, drice_params *out);

/* Worst-case compressed size in BYTES of one chunk of `nsamples` int16 cut into waves of L
 * samples (L <= 0: one wave): 4*(1 + W + sum ceil(25*len/32)). */
DRICE_API size_t drice_chunk_bound_bytes(size_t nsamples, int64_t L);
/* Same for a batch: chunk c holds samples [off[c], off[c+1]). */
DRICE_API size_t drice_batch_bound_bytes(const uint64_t *chunk_sample_off, size_t nchunks, int64_t L);

/* Context: one CUDA device, its streams, scratch and pinned staging.  `device` < 0 uses the
 * current device.  Not thread-safe per context; use one context per host thread.  The device
 * scratch (tickets, look-back words, wave tables) is one set per context: calls may be enqueued on
 * different streams, a call on another stream than the previous one is ordered behind that call's
 * last kernel (so two contexts, not two streams, are what runs an encode next to a decode). */
DRICE_API int  drice_create(drice_ctx **ctx, int device);
DRICE_API void drice_destroy(drice_ctx *ctx);
DRICE_API const char *drice_last_error(const drice_ctx *ctx);   /* ctx may be NULL: creation errors */
DRICE_API int  drice_device(const drice_ctx *ctx);

/* Pre-filter of the following encode / decode calls on this context (reference cd_values[2..],
 * encodeWaveform / decodeWaveform src/deltaRice.c:49-104): `filter_len` taps.  The default
 * (and filter == NULL) is the delta filter [1,-1], fused into the codec kernels; [1] codes the
 * samples as they are; any other filter adds one FIR pass before encode and one recursive
 * pass after decode, with the reference's arithmetic (sums modulo 2^16, C division by f[0]).
 * DRICE_E_PARAM: filter_len < 1 or f[0] == 0; DRICE_E_UNSUPPORTED: > DRICE_MAX_FILTER taps. */
DRICE_API int drice_set_filter(drice_ctx *ctx, const int32_t *filter, int filter_len);

/* Pinned host memory helpers for the host-pointer entry points (optional). */
DRICE_API void *drice_host_alloc(size_t bytes);
DRICE_API void  drice_host_free(void *p);

/* ---- device-pointer entry points: buffers already resident in HBM ------------------
 * d_raw           int16 samples of all chunks back to back (device)
 * chunk_sample_off host array [nchunks+1], cumulative sample offsets (chunk c = [off[c],off[c+1]))
 * M, L            Rice parameter and WaveformLength (L = -1: whole chunk)
 * d_out           device buffer for the concatenated chunk streams, 4-byte aligned
 * out_cap_bytes   its capacity (>= drice_batch_bound_bytes to be always safe)
 * stream          cudaStream_t as void* (NULL = the context's own stream)
 *
 * The *_async form only enqueues work on `stream`: chunk byte offsets ([nchunks+1], u64)
 * are left in DEVICE memory at d_chunk_byte_off and a status word at *d_status (0 = ok).
 * The synchronous form waits, copies the offsets to the host array and checks status.   */
DRICE_API int drice_encode_batch_dev_async(drice_ctx *ctx, const int16_t *d_raw,
                                 const uint64_t *chunk_sample_off, size_t nchunks,
                                 int M, int64_t L, uint32_t *d_out, size_t out_cap_bytes,
                                 uint64_t *d_chunk_byte_off, uint32_t *d_status, void *stream);
DRICE_API int drice_encode_batch_dev(drice_ctx *ctx, const int16_t *d_raw,
                           const uint64_t *chunk_sample_off, size_t nchunks,
                           int M, int64_t L, uint32_t *d_out, size_t out_cap_bytes,
                           uint64_t *chunk_byte_off /* host, [nchunks+1] */, void *stream);

/* Decode: d_comp holds the chunk streams; chunk c occupies bytes [boff[c], boff[c+1]) and
 * must decode to exactly off[c+1]-off[c] samples (checked against the stream's own count).
 * d_out receives samples at the positions given by chunk_sample_off.                   */
DRICE_API int drice_decode_batch_dev_async(drice_ctx *ctx, const uint32_t *d_comp,
                                 const uint64_t *chunk_byte_off, size_t nchunks,
                                 const uint64_t *chunk_sample_off, int M, int64_t L,
                                 int16_t *d_out, uint32_t *d_status, void *stream);
DRICE_API int drice_decode_batch_dev(drice_ctx *ctx, const uint32_t *d_comp,
                           const uint64_t *chunk_byte_off, size_t nchunks,
                           const uint64_t *chunk_sample_off, int M, int64_t L,
                           int16_t *d_out, void *stream);

/* ---- host-pointer entry points: the chunk scheduler --------------------------------
 * Same contracts with HOST buffers.  Work is cut into sub-batches of whole chunks that are
 * staged through pinned buffers and copied asynchronously so H2D, kernels and D2H of
 * neighbouring sub-batches overlap.  Pinned caller buffers (drice_host_alloc) skip the
 * staging memcpy.                                                                      */
DRICE_API int drice_encode_batch_host(drice_ctx *ctx, const int16_t *h_raw,
                            const uint64_t *chunk_sample_off, size_t nchunks,
                            int M, int64_t L, void *h_out, size_t out_cap_bytes,
                            uint64_t *chunk_byte_off /* host, [nchunks+1] */);
DRICE_API int drice_decode_batch_host(drice_ctx *ctx, const void *h_comp,
                            const uint64_t *chunk_byte_off, size_t nchunks,
                            const uint64_t *chunk_sample_off, int M, int64_t L,
                            int16_t *h_out);

/* Reads the leading u32 (total samples) of each chunk of a HOST stream: what a caller
 * needs to build chunk_sample_off for drice_decode_batch_host.                          */
DRICE_API int drice_peek_chunk_samples(const void *h_comp, const uint64_t *chunk_byte_off,
                             size_t nchunks, uint64_t *chunk_samples /* [nchunks] */);

/* Number of kernels the context has launched so far (bench.py's gpu_launches). */
DRICE_API uint64_t drice_launch_count(const drice_ctx *ctx);

/* Per-kernel device timing (bench.py's roofline leg).  While enabled every kernel launch is
 * bracketed by a pair of CUDA events on the stream it is launched on.  drice_timing_read
 * waits for the pending events and returns, per kernel kind, the summed duration in
 * milliseconds and the number of launches since the last reset. */
#define DRICE_KERNEL_ENCODE 0
#define DRICE_KERNEL_LOCATE 1
#define DRICE_KERNEL_PARSE  2
#define DRICE_NUM_KERNELS   3
DRICE_API int drice_timing_enable(drice_ctx *ctx, int on);
DRICE_API int drice_timing_read(drice_ctx *ctx, double *ms /* [DRICE_NUM_KERNELS] */,
                      uint64_t *launches /* [DRICE_NUM_KERNELS] */, int reset);
DRICE_API const char *drice_kernel_name(int kind);

#ifdef __cplusplus
}
#endif
#endif /* DELTARICE_B200_H */
