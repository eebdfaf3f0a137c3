// tcgen05.ld throughput by shape: W warps of one CTA per SM stream TMEM -> registers in a loop (the accumulators are
// never written: contents are irrelevant), cycles per instruction and bytes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_rate tmem_ld_rate.cu && ./tmem_ld_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int SHAPE> __device__ __forceinline__ uint32_t ld(uint32_t a) {
    uint32_t r[16], x = 0;
    if (SHAPE == 0) {        // 32x32b.x16: 32 lanes x 16 columns = 2048 B
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(a));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) x ^= r[i];
    } else if (SHAPE == 1) { // 32x32b.x8: 1024 B
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(a));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) x ^= r[i];
    } else if (SHAPE == 2) { // 16x256b.x1: 16 lanes x 8 columns = 512 B
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]) : "r"(a));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; ++i) x ^= r[i];
    } else if (SHAPE == 3) { // 16x256b.x2: 1024 B
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(a));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) x ^= r[i];
    } else {                 // 16x256b.x4: 2048 B
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(a));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) x ^= r[i];
    }
    return x;
}
// BATCH loads are issued before one wait when BATCHED (the epilogue's pattern), else wait after every load
template <int SHAPE>
__global__ void __launch_bounds__(512) k(int iters, unsigned long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { uint32_t d = (uint32_t)__cvta_generic_to_shared(&tmem_base); asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(d)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t x = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x ^= ld<SHAPE>(base + ((it * 8 + j) * 16) % 448);
    }
    long long t1 = clock64();
    if (x == 0x12345678u) sink[0] = x;
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base));
}
template <int SHAPE> void run(int warps, unsigned long long* cyc, uint32_t* sink, const char* name, int bytes) {
    const int iters = 500;
    for (int r = 0; r < 2; ++r) k<SHAPE><<<148, warps * 32>>>(iters, cyc, sink);
    cudaDeviceSynchronize();
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 8);
    printf("%-14s %2d warps: %6.1f cycles per load+wait per warp, %7.1f B/clk/SM  (%s)\n", name, warps, per, warps * bytes / per, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    unsigned long long* cyc; uint32_t* sink; cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4);
    for (int w : {4, 8, 16}) {
        run<0>(w, cyc, sink, "32x32b.x16", 2048); run<1>(w, cyc, sink, "32x32b.x8", 1024);
        run<2>(w, cyc, sink, "16x256b.x1", 512); run<3>(w, cyc, sink, "16x256b.x2", 1024); run<4>(w, cyc, sink, "16x256b.x4", 2048);
    }
    return 0;
}
