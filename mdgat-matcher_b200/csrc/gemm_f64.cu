// fp64 GEMM  Y = act(scale * X W^T + b) + Res  on the DMMA.8x8x4 tensor path (sm_100a).
//
// Serves every 1x1-conv of the reference (MLP(), /root/reference/models/mdgat.py:34-46, with
// eval-mode BatchNorm folded by the host packer), the stacked q/k/v projection of
// MultiHeadedAttention.forward (mdgat.py:227-232), the merge conv (:237), final_proj (:397),
// the score einsum (:430-431) and the dense logits of dynamic_attention (:201).
//
// Tiling: CTA 64 x 128 outputs, BK = 32, 8 warps (2 x 4), warp tile 32 x 32 = 4 x 4 DMMA tiles.
// Operands are staged with 16-byte cp.async into shared rows padded to 36 doubles
// (36*2 words = 8 mod 32 banks -> the 8x4 fragment read is conflict-free), double-buffered.
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int G_BM = 64, G_BN = 128, G_BK = 32, G_LDS = 36, G_THREADS = 256;
constexpr int G_STAGES = 2;
constexpr size_t G_SMEM = (size_t)G_STAGES * (G_BM + G_BN) * G_LDS * sizeof(double);

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 2) gemm_f64_kernel(GemmParams p) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                                   // [stage][G_BM][G_LDS]
    double* Ws = smem + G_STAGES * G_BM * G_LDS;         // [stage][G_BN][G_LDS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int m0 = blockIdx.x * G_BM, n0 = blockIdx.y * G_BN;
    const int z = blockIdx.z;
    const double* A0 = p.A0 + (long long)z * p.sA;
    const double* A1 = p.A1;
    const double* W = p.W + (long long)z * p.sW;

    const int nk = (p.K + G_BK - 1) / G_BK;

    auto load_stage = [&](int kc, int buf) {
        const int k0 = kc * G_BK;
        // which input segment feeds this K chunk (K0 is a multiple of G_BK when A1 is used)
        const double* Ab = A0;
        int lda = p.lda0, kk = k0, klim = p.K0;
        if (k0 >= p.K0) { Ab = A1; lda = p.lda1; kk = k0 - p.K0; klim = p.K - p.K0; }
        double* as = As + buf * G_BM * G_LDS;
        double* ws = Ws + buf * G_BN * G_LDS;
#pragma unroll
        for (int i = 0; i < (G_BM * (G_BK / 2)) / G_THREADS; ++i) {
            int c = tid + i * G_THREADS;
            int r = c >> 4, kc2 = (c & 15) * 2;
            bool ok = (m0 + r < p.R) && (kk + kc2 < klim);
            const double* src = ok ? Ab + (long long)(m0 + r) * lda + kk + kc2 : Ab;
            cp_async16(as + r * G_LDS + kc2, src, ok);
        }
#pragma unroll
        for (int i = 0; i < (G_BN * (G_BK / 2)) / G_THREADS; ++i) {
            int c = tid + i * G_THREADS;
            int r = c >> 4, kc2 = (c & 15) * 2;
            bool ok = (n0 + r < p.Nout) && (k0 + kc2 < p.K);
            const double* src = ok ? W + (long long)(n0 + r) * p.ldw + k0 + kc2 : W;
            cp_async16(ws + r * G_LDS + kc2, src, ok);
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    load_stage(0, 0);
    cp_async_commit();
    for (int kc = 0; kc < nk; ++kc) {
        if (kc + 1 < nk) {
            load_stage(kc + 1, (kc + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const double* as = As + (kc & 1) * G_BM * G_LDS + (wm * 32 + (lane >> 2)) * G_LDS + (lane & 3);
        const double* ws = Ws + (kc & 1) * G_BN * G_LDS + (wn * 32 + (lane >> 2)) * G_LDS + (lane & 3);
#pragma unroll
        for (int ks = 0; ks < G_BK / 4; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = as[i * 8 * G_LDS + ks * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = ws[j * 8 * G_LDS + ks * 4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

    // ---- epilogue ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + wm * 32 + i * 8 + (lane >> 2);
        if (row >= p.R) continue;
        long long hrow = 0;      // EPI_QKV: row index inside the head-major buffers (before head offset)
        int npts = 0;
        if (EPI == EPI_QKV) {
            if (row < p.rows0) { int b = row / p.n0; npts = p.n0; hrow = (long long)b * HEADS * p.n0 + (row - b * p.n0); }
            else { int r1 = row - p.rows0; int b = r1 / p.n1; npts = p.n1;
                   hrow = (long long)p.rows0 * HEADS + (long long)b * HEADS * p.n1 + (r1 - b * p.n1); }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
            if (col >= p.Nout) continue;
            double y0 = acc[i][j][0] * p.scale, y1 = acc[i][j][1] * p.scale;
            const bool has1 = (col + 1 < p.Nout);
            if (p.bias) { y0 += p.bias[col]; if (has1) y1 += p.bias[col + 1]; }
            if (p.relu) { y0 = fmax(y0, 0.0); y1 = fmax(y1, 0.0); }
            if (EPI == EPI_PLAIN) {
                if (p.Res) {
                    const double* rr = p.Res + (long long)row * p.ldres + col;
                    y0 += rr[0]; if (has1) y1 += rr[1];
                }
                double* yy = p.Y + (long long)z * p.sY + (long long)row * p.ldy + col;
                if (has1 && ((reinterpret_cast<uintptr_t>(yy) & 15) == 0)) {
                    *reinterpret_cast<double2*>(yy) = make_double2(y0, y1);
                } else { yy[0] = y0; if (has1) yy[1] = y1; }
            } else {
                // col in [0,384): which = col/128 (q,k,v); head-major channel c' = h*32 + d
                const int which = col >> 7, c = col & 127, h = c >> 5, d = c & 31;
                double* base = which == 0 ? p.Qh : (which == 1 ? p.Kh : p.Vh);
                const int ld = which == 2 ? LDH_V : LDH_QK;
                double* yy = base + (hrow + (long long)h * npts) * ld + d;
                *reinterpret_cast<double2*>(yy) = make_double2(y0, y1);   // d even, ld even
            }
        }
    }
}

cudaError_t launch_gemm(const GemmParams& p, int epi, int batch, cudaStream_t st) {
    dim3 grid((p.R + G_BM - 1) / G_BM, (p.Nout + G_BN - 1) / G_BN, batch);
    if (p.R <= 0 || p.Nout <= 0 || batch <= 0) return cudaSuccess;
    // per-device attribute; set on every launch (cheap) so multi-device processes stay correct
    cudaError_t e = epi == EPI_PLAIN
        ? cudaFuncSetAttribute(gemm_f64_kernel<EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM)
        : cudaFuncSetAttribute(gemm_f64_kernel<EPI_QKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
    if (e != cudaSuccess) return e;
    if (epi == EPI_PLAIN) gemm_f64_kernel<EPI_PLAIN><<<grid, G_THREADS, G_SMEM, st>>>(p);
    else gemm_f64_kernel<EPI_QKV><<<grid, G_THREADS, G_SMEM, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mdgat
