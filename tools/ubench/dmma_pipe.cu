// Micro-benchmarks that isolate what keeps a DMMA.8x8x4 GEMM inner loop below the issue-loop peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_pipe dmma_pipe.cu && ./dmma_pipe
#include <cstdio>
#include <cuda_runtime.h>
#define DEVINL __device__ __forceinline__
DEVINL void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
DEVINL void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem));
}
constexpr int LDS_ = 36;
// MODE 0: registers only. 1: + LDS fragments each k-step. 2: + __syncthreads per 8 k-steps (x2).
// 3: + cp.async refill of the other stage from global each chunk (latency exposed: wait_group 0).
// 4: same loads but prefetched one chunk ahead (wait_group 1). 5: + epilogue every 4 chunks (scale, bias,
// relu, double2 stores). 6: + residual loads in the epilogue.
template <int MODE, int MI, int NI, int WROWS = 128, int OCC = 2>
__global__ void __launch_bounds__(256, OCC) k(double* out, const double* g, int chunks) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qr = lane >> 2, qc = lane & 3;
    constexpr int ROWS = 64 + WROWS;
    for (int i = tid; i < 2 * ROWS * LDS_; i += 256) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double a[MI], b[NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = 1.0 + i + lane * 1e-6;
#pragma unroll
    for (int j = 0; j < NI; ++j) b[j] = 0.5 - j;
    // warp grid: (8*MI*wm rows) x (8*NI*wn cols); keep it simple: A rows from [0,64), B rows from [64,192)
    const int wm = (MI == 4 && NI == 4) ? (warp >> 2) : 0, wn = (MI == 4 && NI == 4) ? (warp & 3) : warp;
    for (int c = 0; c < chunks; ++c) {
        const double* as = sm + (c & 1) * ROWS * LDS_ + (wm * 32 + qr) * LDS_ + qc;
        const double* ws = sm + (c & 1) * ROWS * LDS_ + (64 + wn * 8 * NI + qr) * LDS_ + qc;
        if (MODE >= 3) {
            double* dst = sm + ((c + 1) & 1) * ROWS * LDS_;
            const double* src = g + ((size_t)blockIdx.x * 4096 + (c & 7) * 512) % (1 << 22);
#pragma unroll
            for (int i = 0; i < (ROWS * 16) / 256; ++i) { int idx = tid + i * 256; int r = idx >> 4, k2 = (idx & 15) * 2; cp_async16(dst + r * LDS_ + k2, src + r * 32 + k2); }
            asm volatile("cp.async.commit_group;\n" ::);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            if (MODE >= 1) {
#pragma unroll
                for (int i = 0; i < MI; ++i) a[i] = as[i * 8 * LDS_ + ks * 4];
#pragma unroll
                for (int j = 0; j < NI; ++j) b[j] = ws[j * 8 * LDS_ + ks * 4];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (MODE == 3) asm volatile("cp.async.wait_group 0;\n" ::);
        if (MODE >= 4) asm volatile("cp.async.wait_group 1;\n" ::);
        if (MODE >= 2) { __syncthreads(); }
        if (MODE >= 5 && (c & 3) == 3) {
            double* y = const_cast<double*>(g) + (1 << 21) + ((size_t)blockIdx.x * 56 + qr) * 132 + wn * 8 * NI + 2 * qc;
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) {
                    double y0 = acc[i][j][0] * 1.0000001 + 0.25, y1 = acc[i][j][1] * 1.0000001 + 0.25;
                    y0 = fmax(y0, 0.0); y1 = fmax(y1, 0.0);
                    double* yy = y + i * 8 * 132 + j * 8;
                    if (MODE >= 6) { y0 += yy[0]; y1 += yy[1]; }
                    *reinterpret_cast<double2*>(yy) = make_double2(y0, y1);
                    acc[i][j][0] = acc[i][j][1] = 0.0;
                }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) out[0] = s;
}
template <int MODE, int MI, int NI, int WROWS = 128, int OCC = 2> void run(const char* name, double* d, const double* g) {
    const int chunks = 2000, grid = 148 * OCC;
    const size_t smem = 2 * (64 + WROWS) * LDS_ * sizeof(double);
    cudaFuncSetAttribute(k<MODE, MI, NI, WROWS, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<MODE, MI, NI, WROWS, OCC>, 256, smem);
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(t0); k<MODE, MI, NI, WROWS, OCC><<<grid, 256, smem>>>(d, g, chunks); cudaEventRecord(t1);
        cudaEventSynchronize(t1); cudaEventElapsedTime(&ms, t0, t1);
    }
    double fl = (double)grid * 8 * chunks * 8 * MI * NI * 512.0;
    printf("%-52s occ %d %8.3f ms  %6.2f TFLOP/s  (%s)\n", name, nb, ms, fl / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double *d, *g; cudaMalloc(&d, 64); cudaMalloc(&g, (size_t)(1 << 23) * 8); cudaMemset(g, 0, (size_t)(1 << 23) * 8);
    run<0, 4, 4>("4x4 regs only", d, g);
    run<1, 4, 4>("4x4 + LDS", d, g);
    run<2, 4, 4>("4x4 + LDS + barrier/chunk", d, g);
    run<3, 4, 4>("4x4 + LDS + barrier + cp.async", d, g);
    run<0, 7, 2>("7x2 regs only", d, g);
    run<1, 7, 2>("7x2 + LDS", d, g);
    run<2, 7, 2>("7x2 + LDS + barrier/chunk", d, g);
    run<3, 7, 2>("7x2 + LDS + barrier + cp.async", d, g);
    run<4, 7, 2>("7x2 + prefetched cp.async (wait 1)", d, g);
    run<5, 7, 2>("7x2 + prefetch + epilogue/4 chunks", d, g);
    run<6, 7, 2>("7x2 + prefetch + epilogue + residual", d, g);
    run<4, 4, 4>("4x4 + prefetched cp.async (wait 1)", d, g);
    run<6, 4, 4>("4x4 + prefetch + epilogue + residual", d, g);
    run<6, 7, 1, 64, 3>("7x1 W64 3 CTAs/SM: prefetch+epilogue+residual", d, g);
    run<6, 7, 1, 64, 2>("7x1 W64 2 CTAs/SM: prefetch+epilogue+residual", d, g);
    run<6, 7, 2, 128, 1>("7x2 W128 1 CTA/SM: prefetch+epilogue+residual", d, g);
    run<4, 7, 1, 64, 3>("7x1 W64 3 CTAs/SM: prefetch only", d, g);
    run<2, 7, 1, 64, 3>("7x1 W64 3 CTAs/SM: LDS+barrier", d, g);
    run<1, 8, 2>("8x2 + LDS", d, g);
    run<1, 4, 2>("4x2 + LDS", d, g);
    return 0;
}
