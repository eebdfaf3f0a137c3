// fp64 GEMM  Y = act(scale * X W^T + b) + Res  on the DMMA.8x8x4 tensor path (sm_100a).
//
// Serves every 1x1-conv of the reference (MLP(), /root/reference/models/mdgat.py:34-46, with
// eval-mode BatchNorm folded by the host packer), the stacked q/k/v projection of
// MultiHeadedAttention.forward (mdgat.py:227-232), the merge conv (:237), final_proj (:397)
// and the score einsum (:430-431).
//
// Scheduling. The row count of the benchmark shape (2*32*512 = 32768 rows) does not divide
// into 64-row tiles evenly over 148 SMs x 2 resident CTAs (1.73 waves -> 14% of the SM-time idle
// in the tail). Instead the rows are cut into one contiguous stripe per resident CTA slot and
// the tile height 8*MI (MI = 4..8 DMMA m-tiles) is picked per launch so that a whole number of
// tiles covers a stripe (32768 rows: 296 stripes of 112 rows = 2 tiles of 56). A CTA walks its
// stripe tile by tile with ONE flat software pipeline over (tile, k-chunk): the first chunk of
// the next tile is already in flight while the epilogue of the current tile runs.
//
// Tiling: CTA tile (8*MI) x 128 outputs, BK = 32, 8 warps side by side along N (warp tile
// 8*MI x 16 = MI x 2 DMMA tiles). Operands are staged with 16-byte cp.async into shared rows
// padded to 36 doubles (36*2 words = 8 mod 32 banks -> the 8x4 fragment read is
// conflict-free), double-buffered.
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int G_BN = 128, G_BK = 32, G_LDS = 36, G_THREADS = 256, G_STAGES = 2;

template <int MI> constexpr size_t gemm_smem_bytes() {
    return (size_t)G_STAGES * (8 * MI + G_BN) * G_LDS * sizeof(double);
}

// WT = true: W is a tile-major blob weight; each (column tile, k chunk) stage is ONE contiguous
// 128 x 36 block moved by a single TMA bulk copy (cp.async.bulk -> UBLKCP) that one thread issues and an
// mbarrier completes. The A rows stay on 16-byte cp.async (they are strided in memory).
template <int MI, int EPI, bool WT>
__global__ void __launch_bounds__(G_THREADS, 2) gemm_f64_kernel(GemmParams p, int rows_per_stripe) {
    constexpr int BM = 8 * MI;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                                   // [stage][BM][G_LDS]
    double* Ws = smem + G_STAGES * BM * G_LDS;           // [stage][G_BN][G_LDS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qr = lane >> 2, qc = lane & 3;
    const int n0 = blockIdx.y * G_BN;
    const int z = blockIdx.z;
    const double* A0 = p.A0 + (long long)z * p.sA;
    const double* A1 = p.A1;
    const double* W = p.W + (long long)z * p.sW;

    const int row_begin = blockIdx.x * rows_per_stripe;
    const int row_end = min(p.R, row_begin + rows_per_stripe);
    if (row_begin >= row_end) return;
    pdl_wait();
    pdl_trigger();
    const int ntiles = (row_end - row_begin + BM - 1) / BM;
    const int nk = (p.K + G_BK - 1) / G_BK;
    const int total = ntiles * nk;
    __shared__ __align__(8) uint64_t wbar[G_STAGES];
    if (WT) {
        if (tid == 0) { mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1); mbar_fence_init(); }
        __syncthreads();
    }

    auto load_stage = [&](int q, int buf) {
        const int t = q / nk, kc = q - t * nk;
        const int m0 = row_begin + t * BM;
        const int k0 = kc * G_BK;
        // which input segment feeds this K chunk (K0 is a multiple of G_BK when A1 is used)
        const double* Ab = A0;
        int lda = p.lda0, kk = k0, klim = p.K0;
        if (k0 >= p.K0) { Ab = A1; lda = p.lda1; kk = k0 - p.K0; klim = p.K - p.K0; }
        double* as = As + buf * BM * G_LDS;
        double* ws = Ws + buf * G_BN * G_LDS;
        for (int c = tid; c < BM * (G_BK / 2); c += G_THREADS) {
            const int r = c >> 4, kc2 = (c & 15) * 2;
            const bool ok = (m0 + r < row_end) && (kk + kc2 < klim);
            const double* src = ok ? Ab + (long long)(m0 + r) * lda + kk + kc2 : Ab;
            cp_async16(as + r * G_LDS + kc2, src, ok);
        }
        if (WT) {
            if (tid == 0) {
                constexpr unsigned bytes = G_BN * G_LDS * sizeof(double);
                mbar_expect_tx(&wbar[buf], bytes);
                bulk_g2s(ws, W + ((size_t)blockIdx.y * nk + kc) * (G_BN * G_LDS), bytes, &wbar[buf]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < (G_BN * (G_BK / 2)) / G_THREADS; ++i) {
                const int c = tid + i * G_THREADS;
                const int r = c >> 4, kc2 = (c & 15) * 2;
                const bool ok = (n0 + r < p.Nout) && (k0 + kc2 < p.K);
                const double* src = ok ? W + (long long)(n0 + r) * p.ldw + k0 + kc2 : W;
                cp_async16(ws + r * G_LDS + kc2, src, ok);
            }
        }
    };

    double acc[MI][2][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    load_stage(0, 0);
    cp_async_commit();
    for (int q = 0; q < total; ++q) {
        if (q + 1 < total) {
            load_stage(q + 1, (q + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        if (WT) mbar_wait(&wbar[q & 1], (unsigned)((q >> 1) & 1));
        __syncthreads();
        const double* as = As + (q & 1) * BM * G_LDS + qr * G_LDS + qc;
        const double* ws = Ws + (q & 1) * G_BN * G_LDS + (warp * 16 + qr) * G_LDS + qc;
#pragma unroll
        for (int ks = 0; ks < G_BK / 4; ++ks) {
            double a[MI], b[2];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = as[i * 8 * G_LDS + ks * 4];
            b[0] = ws[ks * 4];
            b[1] = ws[8 * G_LDS + ks * 4];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                dmma884(acc[i][0][0], acc[i][0][1], a[i], b[0]);
                dmma884(acc[i][1][0], acc[i][1][1], a[i], b[1]);
            }
        }
        __syncthreads();

        const int t = q / nk;
        if (q - t * nk != nk - 1) continue;
        // ---- epilogue of tile t (the first chunk of tile t+1 is already in flight)
        const int m0 = row_begin + t * BM;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int row = m0 + i * 8 + qr;
            if (row < row_end) {
                long long hrow = 0;      // EPI_QKV: row index inside the head-major buffers (before head offset)
                int npts = 0;
                if (EPI == EPI_QKV) {
                    if (row < p.rows0) { const int b = row / p.n0; npts = p.n0; hrow = (long long)b * HEADS * p.n0 + (row - b * p.n0); }
                    else { const int r1 = row - p.rows0; const int b = r1 / p.n1; npts = p.n1;
                           hrow = (long long)p.rows0 * HEADS + (long long)b * HEADS * p.n1 + (r1 - b * p.n1); }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int col = n0 + warp * 16 + j * 8 + 2 * qc;
                    if (col >= p.Nout) continue;
                    // the epilogue competes with the other warps' DMMAs for the FP64 pipe: keep its
                    // FP64 instruction count minimal (no multiply when scale == 1, ReLU on the sign bit)
                    double y0 = acc[i][j][0], y1 = acc[i][j][1];
                    if (p.scale != 1.0) { y0 *= p.scale; y1 *= p.scale; }
                    const bool has1 = (col + 1 < p.Nout);
                    if (p.bias) { y0 += p.bias[col]; if (has1) y1 += p.bias[col + 1]; }
                    if (p.relu) { y0 = __double2hiint(y0) < 0 ? 0.0 : y0; y1 = __double2hiint(y1) < 0 ? 0.0 : y1; }
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
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        }
    }
}

template <int MI, int EPI, bool WT>
static cudaError_t launch_inst(const GemmParams& p, dim3 grid, int rps, cudaStream_t st) {
    const size_t smem = gemm_smem_bytes<MI>();
    // per-device attribute; set on every launch (cheap) so multi-device processes stay correct
    cudaError_t e = cudaFuncSetAttribute(gemm_f64_kernel<MI, EPI, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_pdl(gemm_f64_kernel<MI, EPI, WT>, grid, dim3(G_THREADS), smem, st, p, rps);
}

template <int MI>
static cudaError_t launch_mi(const GemmParams& p, int epi, dim3 grid, int rps, cudaStream_t st) {
    if (epi == EPI_QKV) return p.w_tiled ? launch_inst<MI, EPI_QKV, true>(p, grid, rps, st) : cudaErrorInvalidValue;
    return p.w_tiled ? launch_inst<MI, EPI_PLAIN, true>(p, grid, rps, st) : launch_inst<MI, EPI_PLAIN, false>(p, grid, rps, st);
}

static int g_sm_count[64] = {0};

cudaError_t launch_gemm(const GemmParams& p, int epi, int batch, cudaStream_t st) {
    if (p.R <= 0 || p.Nout <= 0 || batch <= 0) return cudaSuccess;
    int dev = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (dev < 64 && g_sm_count[dev] == 0) {
        int sms = 0;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        g_sm_count[dev] = sms;
    }
    const int sms = dev < 64 ? g_sm_count[dev] : 148;
    const int slots = 2 * sms;                                   // two CTAs of this kernel are resident per SM
    const int ny = (p.Nout + G_BN - 1) / G_BN;
    const long long units = (long long)ny * batch;
    // one stripe per resident slot, but not thinner than 32 rows
    long long nstripes = slots / units;
    if (nstripes < 1) nstripes = 1;
    const long long max_stripes = (p.R + 31) / 32;
    if (nstripes > max_stripes) nstripes = max_stripes;
    int rps = (int)((p.R + nstripes - 1) / nstripes);
    rps = (rps + 7) & ~7;
    nstripes = (p.R + rps - 1) / rps;
    // tile height: least DMMA row-slots per stripe, the taller tile on ties
    int best_mi = 8;
    long long best_cost = -1;
    for (int mi = 8; mi >= 4; --mi) {
        const int bm = 8 * mi;
        const long long tiles = (rps + bm - 1) / bm;
        const long long cost = tiles * bm;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_mi = mi; }
    }
    dim3 grid((unsigned)nstripes, ny, batch);
    switch (best_mi) {
        case 4: e = launch_mi<4>(p, epi, grid, rps, st); break;
        case 5: e = launch_mi<5>(p, epi, grid, rps, st); break;
        case 6: e = launch_mi<6>(p, epi, grid, rps, st); break;
        case 7: e = launch_mi<7>(p, epi, grid, rps, st); break;
        default: e = launch_mi<8>(p, epi, grid, rps, st); break;
    }
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

}  // namespace mdgat
