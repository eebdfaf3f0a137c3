// Stand-alone check of a tcgen05.mma kind::i8 tile (M=128, N=64, K=128) with hand-built descriptors
// and operands in the canonical no-swizzle K-major core-matrix layout, against a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_test umma_i8_test.cu && ./umma_i8_test
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 128;

__device__ __forceinline__ uint64_t make_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;                 // descriptor version (Blackwell)
    return d;                        // layout_type = 0 (no swizzle), base_offset = 0
}

// canonical K-major layout of a (rows x K) int8 tile: core matrix = 8 rows x 16 bytes, 128 B contiguous;
// k-chunks of a row group are adjacent (LBO = 128), row groups follow each other (SBO = K*8)
__host__ __device__ inline int canon(int r, int k) { return (r / 8) * (K * 8) + (k / 16) * 128 + (r % 8) * 16 + (k % 16); }

__global__ void __launch_bounds__(128) kern(const int8_t* A, const int8_t* B, int32_t* D) {
    __shared__ __align__(128) int8_t sA[M * K];
    __shared__ __align__(128) int8_t sB[N * K];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K / 16; i += 128) reinterpret_cast<int4*>(sA)[i] = reinterpret_cast<const int4*>(A)[i];
    for (int i = tid; i < N * K / 16; i += 128) reinterpret_cast<int4*>(sB)[i] = reinterpret_cast<const int4*>(B)[i];
    if (tid == 0) {
        uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(dst));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // generic-proxy smem writes must be visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        // instruction descriptor: D = s32, A = B = signed 8 bit, both K-major, N >> 3, M >> 4
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int kk = 0; kk < K / 32; ++kk) {
            const uint64_t da = make_desc(sA + kk * 256, 128, K * 8);
            const uint64_t db = make_desc(sB + kk * 256, 128, K * 8);
            const uint32_t acc = kk > 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                         "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                         :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(b) : "memory");
    }
    {   // wait for the MMAs
        uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // each warp reads its 32 lanes (rows), 32 columns at a time
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                       "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = (int32_t)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem));
}

int main() {
    std::vector<int8_t> A(M * K), B(N * K), Ac(M * K), Bc(N * K);
    srand(1);
    for (auto& v : A) v = (int8_t)(rand() % 129 - 64);
    for (auto& v : B) v = (int8_t)(rand() % 129 - 64);
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) Ac[canon(r, k)] = A[r * K + k];
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) Bc[canon(r, k)] = B[r * K + k];
    int8_t *dA, *dB; int32_t* dD;
    cudaMalloc(&dA, M * K); cudaMalloc(&dB, N * K); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, Ac.data(), M * K, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bc.data(), N * K, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * N * 4);
    kern<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<int32_t> D(M * N);
    cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
        int32_t s = 0;
        for (int k = 0; k < K; ++k) s += (int32_t)A[i * K + k] * (int32_t)B[j * K + k];
        if (s != D[i * N + j]) { if (bad < 5) printf("mismatch (%d,%d): got %d want %d\n", i, j, D[i * N + j], s); ++bad; }
    }
    printf("mismatches: %ld of %d\n", bad, M * N);
    return bad != 0;
}
