// tcgen05.mma kind::i8 probe of the ISSUE PATTERNS the Ozaki GEMM / attention kernels use, one CTA per SM,
// operands resident in shared memory (no loads in the timed loop):
//   rate   : 16 back-to-back MMAs of M=128, K=32 for N in {32..256} (incl. the non-power-of-two widths)
//   stacked: one Ozaki "unit" = KS k-steps x S MMAs, activation plane s against weight planes 0..S-1-s
//            (N = (S-s)*32, written 32*s columns into the accumulator set), alternating accumulator sets
//   +ld    : the same while 8 other warps stream tcgen05.ld.32x32b.x16 over the other accumulator set
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_pattern umma_i8_pattern.cu && ./umma_i8_pattern
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t make_desc(const void* smem, uint32_t lbo, uint32_t sbo) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, bool acc) {
    if (acc) asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" :: "r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
    else asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" :: "r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(b) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
}
constexpr uint32_t IBASE = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);

// MODE 0: rate of a single width N (16 MMAs per iteration, 4 k-steps x 4 accumulator offsets)
// MODE 1: stacked Ozaki unit, S planes, KC = 128 (4 k-steps), M = 128; XK = K extent of the smem tile (bytes per row)
// LD: 1 = the 8 "epilogue" warps hammer tcgen05.ld on the other accumulator set meanwhile
template <int MODE, int N, int S, int LD>
__global__ void __launch_bounds__(320) k(int iters, unsigned long long* cycles) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int XT = 128 * 128, WT = 32 * 128;
    for (int i = tid; i < (S * XT + S * WT + 256 * 128) / 4; i += 320) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u * (i & 3);
    if (tid == 0) { uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); stop = 0; }
    if (warp == 8) { uint32_t d = (uint32_t)__cvta_generic_to_shared(&tmem_base); asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(d)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    if (warp == 8) {
        if ((tid & 31) == 0) {
            const uint64_t xd0 = make_desc(sm, 128, 1024), wd0 = make_desc(sm + S * XT, 128, 1024);
            long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                if (MODE == 0) {
                    const uint32_t idesc = IBASE | ((uint32_t)(N >> 3) << 17);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int kk = j & 3;
                        mma_i8(tmem + (N <= 128 ? (j >> 2) * N % 512 : 0), xd0 + (uint64_t)((kk * 256) >> 4), wd0 + (uint64_t)((kk * 256) >> 4), idesc, true);
                    }
                } else {
                    const uint32_t dbase = tmem + (it & 1) * (S * 32);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            const uint32_t idesc = IBASE | ((uint32_t)(((S - s) * 32) >> 3) << 17);
                            mma_i8(dbase + s * 32, xd0 + (uint64_t)((s * XT + kk * 256) >> 4), wd0 + (uint64_t)((kk * 256) >> 4), idesc, s > 0 || kk > 0);
                        }
                }
            }
            commit(&bar);
            wait(&bar, 0);
            long long t1 = clock64();
            if (blockIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
            stop = 1;
        }
    } else if (warp < 8 && LD) {
        uint32_t sink = 0;
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 16;
        while (!stop) {
#pragma unroll
            for (int dd = 0; dd < S; ++dd) {
                uint32_t r[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),
                               "=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
                             : "r"(lane_addr + (dd * 32) % 224));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) sink ^= r[j];
            }
        }
        if (sink == 0x12345678u) cycles[1] = sink;
    }
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

template <int MODE, int N, int S, int LD> void run(unsigned long long* d, const char* what) {
    const int iters = 1000; const size_t smem = (size_t)S * (128 * 128 + 32 * 128) + 256 * 128 + 1024;
    cudaFuncSetAttribute(k<MODE, N, S, LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1); float ms = 0;
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(t0); k<MODE, N, S, LD><<<148, 320, smem>>>(iters, d); cudaEventRecord(t1); cudaEventSynchronize(t1); cudaEventElapsedTime(&ms, t0, t1); }
    unsigned long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double macs = MODE == 0 ? 16.0 * 128 * N * 32 : 4.0 * 128 * 32 * (32.0 * S * (S + 1) / 2);
    printf("%-28s %8.1f cycles/%s  %6.0f TOP/s chip  %.3f ms (%s)\n", what, (double)c / iters / (MODE == 0 ? 16 : 1), MODE == 0 ? "MMA" : "unit",
           148.0 * iters * 2.0 * macs / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 16);
    run<0, 32, 1, 0>(d, "rate N=32"); run<0, 64, 1, 0>(d, "rate N=64"); run<0, 96, 1, 0>(d, "rate N=96"); run<0, 128, 1, 0>(d, "rate N=128");
    run<0, 160, 1, 0>(d, "rate N=160"); run<0, 192, 1, 0>(d, "rate N=192"); run<0, 224, 1, 0>(d, "rate N=224"); run<0, 256, 1, 0>(d, "rate N=256");
    run<1, 0, 7, 0>(d, "stacked unit S=7"); run<1, 0, 6, 0>(d, "stacked unit S=6"); run<1, 0, 4, 0>(d, "stacked unit S=4");
    run<1, 0, 7, 1>(d, "stacked unit S=7 + tmem ld");
    return 0;
}
