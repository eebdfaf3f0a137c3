// Global-store throughput of the access patterns a tcgen05 epilogue can produce, one 512-thread CTA per SM (as in the
// Ozaki GEMM), rows of 132 doubles (1056 B) like the activation buffers:
//   coalesced : every warp instruction writes 512 contiguous bytes
//   quad64    : MMA C-fragment order, a quad writes 64 contiguous bytes of a row, a warp 8 rows (16x256b TMEM shape)
//   row16     : thread = row, 16 bytes per thread per instruction, 32 rows per warp instruction (32x32b TMEM shape)
//   bulk256   : rows staged in shared memory, one cp.async.bulk shared->global of 256 B per row
// Working sets: 64 MB (fits the L2) and 1 GB.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu && ./store_patterns
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int LD = 132;          // doubles per row
constexpr int COLS = 128;

// each CTA walks row tiles of 128 rows (tile = blockIdx.x, += gridDim.x); per tile 4 units of 32 columns
template <int MODE>
__global__ void __launch_bounds__(512) k(double* Y, int row_tiles, int reps) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int rep = 0; rep < reps; ++rep)
    for (int tile = blockIdx.x; tile < row_tiles; tile += gridDim.x) {
        double* base = Y + (size_t)tile * 128 * LD;
        for (int unit = 0; unit < COLS / 32; ++unit) {
            const double v = (double)(tile + unit + rep);
            if (MODE == 0) {
                // 128 rows x 32 cols: 16 warps, warp w writes rows 8w..8w+7: lane -> row 8w + lane/4... keep 512 B contiguous:
                // a row's 32 columns = 256 B; a warp instruction covers 2 rows x 256 B
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = warp * 8 + i * 2 + (lane >> 4), col = unit * 32 + (lane & 15) * 2;
                    *reinterpret_cast<double2*>(base + (size_t)row * LD + col) = make_double2(v, v);
                }
            } else if (MODE == 1) {
                const int quarter = warp & 3, cg = warp >> 2;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = quarter * 32 + (lane >> 2) + 8 * i, col = unit * 32 + cg * 8 + (lane & 3) * 2;
                    *reinterpret_cast<double2*>(base + (size_t)row * LD + col) = make_double2(v, v);
                }
            } else if (MODE == 2) {
                // 16 warps: warp w -> rows 32*(w%4) + lane, columns 8*(w/4).. (4 double2 per thread)
                const int row = (warp & 3) * 32 + lane, col = unit * 32 + (warp >> 2) * 8;
#pragma unroll
                for (int i = 0; i < 4; ++i) *reinterpret_cast<double2*>(base + (size_t)row * LD + col + 2 * i) = make_double2(v, v);
            } else {
                // stage 128 rows x 256 B in shared memory (row stride 272 B), then one bulk copy per row issued by 128 threads
                double* st = reinterpret_cast<double*>(sm);
                const int row = (warp & 3) * 32 + lane, c0 = (warp >> 2) * 8;
#pragma unroll
                for (int i = 0; i < 4; ++i) *reinterpret_cast<double2*>(st + (size_t)row * 34 + c0 + 2 * i) = make_double2(v, v);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (tid < 128) {
                    const uint32_t s = (uint32_t)__cvta_generic_to_shared(st + (size_t)tid * 34);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;" :: "l"(base + (size_t)tid * LD + unit * 32), "r"(s) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                __syncthreads();
            }
        }
    }
}

template <int MODE> void run(double* Y, size_t bytes, const char* what) {
    const int row_tiles = (int)(bytes / (128 * LD * 8));
    const int reps = bytes < (200u << 20) ? 8 : 1;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 34 * 8);
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1); float ms = 0;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(t0); k<MODE><<<148, 512, 128 * 34 * 8>>>(Y, row_tiles, reps); cudaEventRecord(t1); cudaEventSynchronize(t1); cudaEventElapsedTime(&ms, t0, t1); }
    const double payload = (double)row_tiles * 128 * COLS * 8 * reps;
    printf("%-10s %5zu MB: %8.3f ms  %6.2f TB/s payload  %5.1f B/clk/SM @1.965GHz (%s)\n", what, bytes >> 20, ms, payload / (ms * 1e-3) / 1e12,
           payload / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double* Y; const size_t big = 1u << 30;
    cudaMalloc(&Y, big); cudaMemset(Y, 0, big);
    for (size_t bytes : {(size_t)64 << 20, big}) {
        run<0>(Y, bytes, "coalesced"); run<1>(Y, bytes, "quad64"); run<2>(Y, bytes, "row16"); run<3>(Y, bytes, "bulk256");
    }
    return 0;
}
