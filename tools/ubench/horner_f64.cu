// Cost of the Ozaki epilogue's float64 recombination in isolation: per "unit" a thread turns 4 merged int32 groups x 8
// outputs into 8 doubles (magic-constant conversion + Horner) and applies the scale FMA -- 32 LEA.HI + 32 DADD + 32 DFMA.
// Measured for 1..4 warps per SM sub-partition; variant B converts with I2F.F64 instead of the magic constant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o horner_f64 horner_f64.cu && ./horner_f64
#include <cstdio>
#include <cuda_runtime.h>
template <int VARIANT>
__global__ void __launch_bounds__(512) k(int iters, double* out, unsigned long long* cycles, int seed) {
    int m[4][8];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 8; ++j) m[g][j] = (threadIdx.x * 131 + g * 17 + j * 7 + seed) & 0xfffff;
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    const double MAGIC = 6755399441055744.0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double h;
            if (VARIANT == 0) {
                auto f = [&](int v) { return __hiloint2double(0x43380000 + (v >> 31), v) - MAGIC; };
                h = f(m[3][j]);
                h = fma(h, 0.00006103515625, f(m[2][j]));
                h = fma(h, 0.00006103515625, f(m[1][j]));
                h = fma(h, 0.00006103515625, f(m[0][j]));
            } else {
                h = (double)m[3][j];
                h = fma(h, 0.00006103515625, (double)m[2][j]);
                h = fma(h, 0.00006103515625, (double)m[1][j]);
                h = fma(h, 0.00006103515625, (double)m[0][j]);
            }
            acc[j] = fma(h, 1.0009765625, acc[j]);
#pragma unroll
            for (int g = 0; g < 4; ++g) m[g][j] += it | 1;
        }
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
}
template <int VARIANT> void run(int threads, double* out, unsigned long long* cyc) {
    const int iters = 2000;
    for (int r = 0; r < 2; ++r) k<VARIANT><<<148, threads>>>(iters, out, cyc, r);
    cudaDeviceSynchronize();
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("variant %d  %2d warps/SM (%d per sub-partition): %7.1f cycles per unit-iteration of a warp set  (%s)\n", VARIANT, threads / 32, threads / 128,
           (double)c / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double* out; unsigned long long* cyc; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 8);
    for (int t : {128, 256, 512}) run<0>(t, out, cyc);
    for (int t : {128, 256, 512}) run<1>(t, out, cyc);
    return 0;
}
