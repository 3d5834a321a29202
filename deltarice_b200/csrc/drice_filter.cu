// drice_filter.cu — the GENERIC pre-filter of the reference (cd_values[2..], any taps other than
// the delta filter [1,-1]) for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   encodeWaveform generic branch   src/deltaRice.c:64-74    -> prefilter_kernel  (FIR)
//   decodeWaveform generic branch   src/deltaRice.c:91-102   -> postfilter_kernel (recursion + /f[0])
//
// The delta filter never comes here: it is fused into encode_tile_kernel / parse_kernel.  A generic
// filter costs one extra pass over the raw samples on each side: encode = prefilter_kernel (raw ->
// scratch) + the encode kernels with their delta switched off; decode = parse_kernel with its
// inverse delta switched off + postfilter_kernel in place.  Filter [1] needs neither pass.
// Both follow the reference's arithmetic exactly: sums and the recursion run modulo 2^16 (the
// reference accumulates in a `short`), the division is C's truncating int division.
#include "drice_kernels.cuh"

namespace drice {

namespace {

constexpr int kFiltThreads = 256;
constexpr int kFiltPerThread = 8;

// one CTA column per chunk (blockIdx.x), grid-stride over the chunk's samples in y
__global__ void __launch_bounds__(kFiltThreads) prefilter_kernel(const FilterParams p, const int16_t *__restrict__ in,
                                                                  int16_t *__restrict__ out)
{
    const uint32_t c = blockIdx.x;
    const uint64_t cb = p.chunk_sample_off[c], ce = p.chunk_sample_off[c + 1];
    const uint64_t total = ce - cb;
    const uint64_t Lw = p.L ? (uint64_t)p.L : total;
    const uint64_t stride = (uint64_t)gridDim.y * kFiltThreads * kFiltPerThread;
    for (uint64_t base = (uint64_t)blockIdx.y * kFiltThreads * kFiltPerThread; base < total; base += stride) {
#pragma unroll
        for (int e = 0; e < kFiltPerThread; ++e) {
            const uint64_t s = base + (uint64_t)e * kFiltThreads + threadIdx.x;     // sample within the chunk
            if (s >= total) break;
            const uint64_t i = s % Lw;                                              // sample within its wave
            const int16_t *x = in + cb + s;
            uint32_t acc = (uint32_t)((int)x[0] * p.f[0]);
            for (int j = 1; j < p.flen && (uint64_t)j <= i; ++j) acc += (uint32_t)((int)x[-j] * p.f[j]);
            out[cb + s] = (int16_t)(uint16_t)acc;
        }
    }
}

// one THREAD per wave (the recursion is serial per wave), in place
__global__ void __launch_bounds__(kFiltThreads) postfilter_kernel(const FilterParams p, int16_t *data, const uint32_t *chunk_wave_off)
{
    const uint32_t c = blockIdx.x;
    const uint64_t cb = p.chunk_sample_off[c], ce = p.chunk_sample_off[c + 1];
    const uint64_t total = ce - cb;
    if (total == 0) return;
    const uint64_t Lw = p.L ? (uint64_t)p.L : total;
    const uint64_t W = (total + Lw - 1) / Lw;
    (void)chunk_wave_off;
    for (uint64_t w = (uint64_t)blockIdx.y * kFiltThreads + threadIdx.x; w < W; w += (uint64_t)gridDim.y * kFiltThreads) {
        int16_t *y = data + cb + w * Lw;
        const uint64_t n = (total - w * Lw) < Lw ? (total - w * Lw) : Lw;
        const int f0 = p.f[0];
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t t = (uint16_t)y[i];
            for (int j = 1; j < p.flen && (uint64_t)j <= i; ++j) t -= (uint32_t)((int)y[i - (uint64_t)j] * p.f[j]);
            y[i] = (int16_t)((int)(int16_t)(uint16_t)t / f0);
        }
    }
}

}  // namespace

int launch_prefilter(const FilterParams &p, const int16_t *in, int16_t *out, uint64_t max_chunk_samples, cudaStream_t st)
{
    if (p.nchunks == 0 || max_chunk_samples == 0) return 0;
    uint64_t gy = (max_chunk_samples + (uint64_t)kFiltThreads * kFiltPerThread - 1) / ((uint64_t)kFiltThreads * kFiltPerThread);
    if (gy > 1024) gy = 1024;
    prefilter_kernel<<<dim3(p.nchunks, (unsigned)gy), kFiltThreads, 0, st>>>(p, in, out);
    return 1;
}

int launch_postfilter(const FilterParams &p, int16_t *data, uint64_t max_chunk_waves, cudaStream_t st)
{
    if (p.nchunks == 0 || max_chunk_waves == 0) return 0;
    uint64_t gy = (max_chunk_waves + kFiltThreads - 1) / kFiltThreads;
    if (gy > 1024) gy = 1024;
    postfilter_kernel<<<dim3(p.nchunks, (unsigned)gy), kFiltThreads, 0, st>>>(p, data, nullptr);
    return 1;
}

}  // namespace drice
