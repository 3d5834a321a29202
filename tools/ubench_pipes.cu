// ubench_pipes.cu — issue-rate microbenchmark for the integer instructions the codec is built from.
// Prints warp-instructions per cycle per SM sub-partition (SMSP) for each instruction and for mixes,
// to decide which work goes to the ALU pipe, the FMA pipe (IMAD family) or the LSU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/ubench_pipes tools/ubench_pipes.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 8;      // independent dependency chains per thread
constexpr int kIters = 512;

#define OP_LOP3(x, y)  asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(c0))
#define OP_SHF(x, y)   asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(c0))
#define OP_SHFR(x, y)  asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(x))
#define OP_SHL(x, y)   asm volatile("shl.b32 %0, %0, %1;" : "+r"(x) : "r"(c0))
#define OP_PRMT(x, y)  asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(c0))
#define OP_ADD(x, y)   asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define OP_ADD3(x, y)  asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(y), "r"(c0))
#define OP_VADD2(x, y) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x) : "r"(y))
#define OP_IMAD(x, y)  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(c0))
#define OP_IMADI(x, y) asm volatile("mad.lo.u32 %0, %0, 0xFFFF0001, %1;" : "+r"(x) : "r"(y))
#define OP_IHI(x, y)   asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(c0))
#define OP_IHII(x, y)  asm volatile("mad.hi.u32 %0, %0, 65536, %1;" : "+r"(x) : "r"(y))
#define OP_WIDE(x, y)  asm volatile("{.reg .u64 t; .reg .u32 lo, hi; mul.wide.u32 t, %0, %1; mov.b64 {lo, hi}, t; xor.b32 %0, lo, hi;}" : "+r"(x) : "r"(y))
#define OP_WIDEONLY(x, y) asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, %1; mov.b64 {%0, %1}, t;}" : "+r"(x), "+r"(y))
#define OP_SETP(x, y)  asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; @p add.u32 %0, %0, %2;}" : "+r"(x) : "r"(y), "r"(c0))
#define OP_LDS(x, y)   asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(sa))
#define OP_LDSDEP(x, y) asm volatile("{.reg .u32 a; and.b32 a, %0, 0x7c; add.u32 a, a, %1; ld.shared.u32 %0, [a];}" : "+r"(x) : "r"(sa))
#define OP_STS(x, y)   asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa), "r"(x) : "memory")
#define OP_SHFL(x, y)  asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(x))

#define BODY1(OP)                                                                                  \
    _Pragma("unroll") for (int c = 0; c < kChains; ++c) { OP(v[c], w[c]); }
#define BODY2(OPA, OPB)                                                                            \
    _Pragma("unroll") for (int c = 0; c < kChains; ++c) { OPA(v[c], w[c]); OPB(w[c], v[c]); }
#define BODY3(OPA, OPB, OPC)                                                                       \
    _Pragma("unroll") for (int c = 0; c < kChains; ++c) { OPA(v[c], w[c]); OPB(w[c], v[c]); OPC(v[c], w[c]); }

#define KERNEL(NAME, BODY)                                                                         \
    __global__ void __launch_bounds__(1024) NAME(uint32_t *out, unsigned long long *cyc, uint32_t c0) \
    {                                                                                              \
        __shared__ uint32_t sm[2048];                                                              \
        uint32_t v[kChains], w[kChains];                                                           \
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4 + (threadIdx.x >> 5) * 128; \
        sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 1024] = c0;                                \
        for (int c = 0; c < kChains; ++c) { v[c] = threadIdx.x * 2654435761u + c; w[c] = c0 + c * 40503u + threadIdx.x; } \
        __syncthreads();                                                                           \
        const long long t0 = clock64();                                                            \
        for (int i = 0; i < kIters; ++i) { BODY }                                                  \
        const long long t1 = clock64();                                                            \
        uint32_t s = 0;                                                                            \
        for (int c = 0; c < kChains; ++c) s += v[c] ^ w[c];                                        \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                            \
        if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);                     \
    }

KERNEL(k_lop3, BODY1(OP_LOP3))
KERNEL(k_shf, BODY1(OP_SHF))
KERNEL(k_shfr, BODY1(OP_SHFR))
KERNEL(k_shl, BODY1(OP_SHL))
KERNEL(k_prmt, BODY1(OP_PRMT))
KERNEL(k_add, BODY1(OP_ADD))
KERNEL(k_add3, BODY1(OP_ADD3))
KERNEL(k_vadd2, BODY1(OP_VADD2))
KERNEL(k_imad, BODY1(OP_IMAD))
KERNEL(k_imadi, BODY1(OP_IMADI))
KERNEL(k_ihi, BODY1(OP_IHI))
KERNEL(k_ihii, BODY1(OP_IHII))
KERNEL(k_wide, BODY1(OP_WIDE))
KERNEL(k_wideonly, BODY1(OP_WIDEONLY))
KERNEL(k_setp, BODY1(OP_SETP))
KERNEL(k_lds, BODY1(OP_LDS))
KERNEL(k_ldsdep, BODY1(OP_LDSDEP))
KERNEL(k_sts, BODY1(OP_STS))
KERNEL(k_shfl, BODY1(OP_SHFL))
KERNEL(k_lop3_imad, BODY2(OP_LOP3, OP_IMAD))
KERNEL(k_lop3_shf, BODY2(OP_LOP3, OP_SHF))
KERNEL(k_lop3_prmt, BODY2(OP_LOP3, OP_PRMT))
KERNEL(k_lop3_ihi, BODY2(OP_LOP3, OP_IHI))
KERNEL(k_lop3_wideonly, BODY2(OP_LOP3, OP_WIDEONLY))
KERNEL(k_lop3_vadd2, BODY2(OP_LOP3, OP_VADD2))
KERNEL(k_imad_vadd2, BODY2(OP_IMAD, OP_VADD2))
KERNEL(k_imad_ihi, BODY2(OP_IMAD, OP_IHI))
KERNEL(k_lop3_sts, BODY2(OP_LOP3, OP_STS))
KERNEL(k_lop3_imad_sts, BODY3(OP_LOP3, OP_IMAD, OP_STS))
KERNEL(k_lop3_imad_lop3, BODY3(OP_LOP3, OP_IMAD, OP_LOP3))
KERNEL(k_imad_lop3_imad, BODY3(OP_IMAD, OP_LOP3, OP_IMAD))

typedef void (*kern_t)(uint32_t *, unsigned long long *, uint32_t);

static void run(const char *name, kern_t k, int ops_per_chain_step, int nblk, int nthr, uint32_t *out, unsigned long long *cyc)
{
    k<<<nblk, nthr>>>(out, cyc, 7);
    cudaDeviceSynchronize();
    k<<<nblk, nthr>>>(out, cyc, 7);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-18s ERROR %s\n", name, cudaGetErrorString(e)); return; }
    unsigned long long h[1024];
    cudaMemcpy(h, cyc, sizeof(unsigned long long) * nblk, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < nblk; ++i) if ((double)h[i] > mx) mx = (double)h[i];
    const double warp_instr = (double)kIters * kChains * ops_per_chain_step * (nthr / 32);   // per CTA = per SM
    printf("%-18s warps/SM %2d  cycles %9.0f  instr/cycle/SMSP %.3f\n", name, nthr / 32, mx, warp_instr / mx / 4.0);
}

int main()
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, sizeof(uint32_t) * 1024 * 1024);
    cudaMalloc(&cyc, sizeof(unsigned long long) * 1024);
    printf("SMs %d\n", nsm);
    for (int nthr : {1024, 256}) {
#define R(NAME, N) run(#NAME, NAME, N, nsm, nthr, out, cyc)
        R(k_lop3, 1); R(k_shf, 1); R(k_shfr, 1); R(k_shl, 1); R(k_prmt, 1); R(k_add3, 1); R(k_vadd2, 1);
        R(k_imad, 1); R(k_imadi, 1); R(k_ihi, 1); R(k_wide, 2); R(k_wideonly, 1); R(k_setp, 2);
        R(k_lds, 1); R(k_ldsdep, 3); R(k_sts, 1); R(k_shfl, 1);
        R(k_lop3_imad, 2); R(k_lop3_shf, 2); R(k_lop3_prmt, 2); R(k_lop3_ihi, 2); R(k_lop3_wideonly, 2); R(k_lop3_vadd2, 2);
        R(k_imad_vadd2, 2); R(k_imad_ihi, 2); R(k_lop3_sts, 2); R(k_lop3_imad_sts, 3); R(k_lop3_imad_lop3, 3); R(k_imad_lop3_imad, 3);
    }
    return 0;
}
