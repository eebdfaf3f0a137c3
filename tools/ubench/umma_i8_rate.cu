// tcgen05.mma kind::i8 issue-rate probe: one CTA per SM, one thread issues `iters` x 16 MMAs (M=128, K=32 each)
// on operands resident in shared memory; layouts: no-swizzle interleaved vs 128-byte swizzle; N = 32..256.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t make_desc(const void* smem, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
template <int N, int SWZ>
__global__ void __launch_bounds__(128) k(int iters, unsigned long long* cycles) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u * (i & 3);
    if (tid == 0) { uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) { uint32_t d = (uint32_t)__cvta_generic_to_shared(&tmem_base); asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(d)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // A tile: 128 rows x 128 B at sm; B tile: N rows x 128 B at sm + 16 KB
        const uint64_t da0 = SWZ ? make_desc(sm, 16, 1024, 2) : make_desc(sm, 128, 1024, 0);
        const uint64_t db0 = SWZ ? make_desc(sm + 16384, 16, 1024, 2) : make_desc(sm + 16384, 128, 1024, 0);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int kk = j & 3;
                const uint64_t da = da0 + (uint64_t)((SWZ ? kk * 32 : kk * 256) >> 4);
                const uint64_t db = db0 + (uint64_t)((SWZ ? kk * 32 : kk * 256) >> 4);
                const uint32_t d = tmem + (N <= 128 ? (j >> 2) * N % 512 : 0);
                asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" :: "r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
            }
        }
        uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(b) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}
template <int N, int SWZ> void run(unsigned long long* d) {
    const int iters = 2000; const size_t smem = (128 + 256) * 128 + 1024;
    cudaFuncSetAttribute(k<N, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1); float ms = 0;
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(t0); k<N, SWZ><<<148, 128, smem>>>(iters, d); cudaEventRecord(t1); cudaEventSynchronize(t1); cudaEventElapsedTime(&ms, t0, t1); }
    unsigned long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 16;
    printf("N=%3d %s: %7.1f cycles/MMA   %6.0f TOP/s chip (%s)\n", N, SWZ ? "swizzle128" : "interleave", c / n,
           148.0 * n * 2.0 * 128 * N * 32 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 8);
    run<32, 0>(d); run<64, 0>(d); run<128, 0>(d); run<256, 0>(d);
    run<32, 1>(d); run<64, 1>(d); run<128, 1>(d); run<256, 1>(d);
    return 0;
}
