// float64-faithful GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
// TMEM, operands staged by TMA bulk copies) -- the Ozaki splitting scheme.
//
// tcgen05 has no f64 MMA kind and the FP64 pipe (DMMA/DFMA) tops out at 37 TFLOP/s on B200, while
// the int8 tensor pipe does 4.5 POP/s. Every operand row is written as
//       x = 2^e * sum_{s=1..S} d_s * 2^(1-7s)  (+ remainder < 2^(e-7S)),   d_s integer, |d_s| <= 64,
// with e the exponent of the row maximum (slice_rows_kernel; all steps are exact in float64). Then
//       X W^T = 2^(e_r + f_n - 12) * sum_dd 2^(-7 dd) * [ sum_{s+t-2=dd} D_s G_t^T ]
// where every bracket is an INTEGER matrix product that kind::i8 accumulates exactly in int32
// (|sum| <= 7 * 256 * 64 * 64 < 2^23). The S diagonals dd = 0..S-1 (S(S+1)/2 slice products; the
// dropped ones are below 2^(-7S) relative to the row scales) live side by side in TMEM and are
// recombined in float64 by a Horner pass in the epilogue, which also applies bias, ReLU, the
// residual and, for the q/k/v projection, the head-major scatter. With S = 7 the result agrees
// with a float64 dot product to a few 1e-14 relative to |x|_max |w|_max sqrt(K).
//
// Replaces, for the per-layer projections of the GNN (q/k/v, MLP 256->256, MLP 256->128;
// /root/reference/models/mdgat.py:227-232, 248), the DMMA GEMM of gemm_f64.cu.
//
// Layouts. A slice tile is (128 rows x 128 k) int8 for X and (64 x 128) for W, stored in the
// canonical no-swizzle K-major core-matrix order the UMMA shared-memory descriptor expects
// (8 rows x 16 bytes contiguous, k-chunks adjacent: LBO = 128 B, row groups SBO = 1024 B), so a tile
// is one contiguous block in global memory and one cp.async.bulk brings it in.
//   Xs[row_tile][k_chunk][slice][128*128]      Ws[col_tile][k_chunk][slice][64*128]
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int OZ_BM = 128, OZ_BN = 64, OZ_KC = 128;
constexpr int OZ_XTILE = OZ_BM * OZ_KC, OZ_WTILE = OZ_BN * OZ_KC;       // bytes per slice tile
constexpr int OZ_THREADS = 128;

DEVINL int oz_canon(int r, int k) { return (r >> 3) * (OZ_KC * 8) + (k >> 4) * 128 + (r & 7) * 16 + (k & 15); }

DEVINL double pow2d(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }   // -1022 <= e <= 1023

// ---------------------------------------------------------------------------------------------------
// Slicing: one thread = 16 consecutive k of one row (8 lanes per 128-wide chunk row). Input = concat of
// up to two row-major float64 buffers (K0 + K1 columns, both multiples of 128).
// ---------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256)
slice_rows_kernel(const double* __restrict__ A0, int ld0, int K0, const double* __restrict__ A1, int ld1, int K1,
                  int R, int8_t* __restrict__ Xs, double* __restrict__ rowscale) {
    const int K = K0 + K1, tpr = K / 16;                 // threads per row
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = gid / tpr;
    const int kt = (int)(gid - r * tpr);                 // which 16-wide k group
    const int Rpad = ((R + OZ_BM - 1) / OZ_BM) * OZ_BM;
    if (r >= Rpad) return;                               // warp-uniform: tpr divides 32 or is a multiple of it
    const int k0 = kt * 16;
    double x[16];
    if (r < R) {
        const double* src = k0 < K0 ? A0 + r * ld0 + k0 : A1 + r * ld1 + (k0 - K0);
#pragma unroll
        for (int i = 0; i < 16; i += 2) { const double2 v = *reinterpret_cast<const double2*>(src + i); x[i] = v.x; x[i + 1] = v.y; }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = 0.0;
    }
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) mx = fmax(mx, fabs(x[i]));
    // row maximum over the tpr lanes of this row (tpr = 8 or 16: a power of two <= 32, lanes contiguous)
    for (int o = 1; o < tpr && o < 32; o <<= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);                         // mx = m * 2^e, m in [0.5, 1)  ->  |x| < 2^e
    e = max(-900, min(900, e));
    if (kt == 0) rowscale[r] = pow2d(e - 12);
    // digits: t = x * 2^(6-e); d = rint(t); t = (t - d) * 2^7; ...
    const double sc = pow2d(6 - e);
    const int tile = (int)(r / OZ_BM), rr = (int)(r % OZ_BM);
    const int kchunk = k0 / OZ_KC, kk = k0 % OZ_KC;
    const int nkc = K / OZ_KC;
    int8_t* base = Xs + ((size_t)(tile * nkc + kchunk) * S) * OZ_XTILE + oz_canon(rr, kk);
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] *= sc;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double d = rint(x[i]);
            x[i] = (x[i] - d) * 128.0;
            w[i >> 2] |= (uint32_t)(uint8_t)(int8_t)(int)d << (8 * (i & 3));
        }
        *reinterpret_cast<uint4*>(base + (size_t)s * OZ_XTILE) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---------------------------------------------------------------------------------------------------
// tcgen05 helpers (raw PTX)
// ---------------------------------------------------------------------------------------------------
DEVINL uint64_t umma_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
DEVINL void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
DEVINL void umma_commit(uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(b) : "memory");
}
DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                   "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct OzParams {
    const int8_t* Xs; const double* rowscale;          // sliced activations (R padded to 128) + 2^(e_r - 12)
    const int8_t* Ws; const double* colscale;          // sliced weights + 2^(f_n)
    const double* bias;                                // [Nout] or null
    const double* Res; int ldres;                      // residual (may alias Y) or null
    double* Y; int ldy;
    int R, Nout, K;                                    // K multiple of 128, Nout multiple of 64
    int relu;
    int epi;                                           // EPI_PLAIN / EPI_QKV
    double *Qh, *Kh, *Vh; int rows0, n0, n1;
    int col_tiles_per_cta;                             // blockIdx.y selects a group of column tiles
};

// One CTA: one 128-row tile x a group of 64-column tiles. Thread 0 stages operands (bulk copies) and issues
// the MMAs; all four warps run the epilogue, warp w owning TMEM lanes (= rows) 32w..32w+31.
template <int S>
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm_kernel(OzParams p) {
    extern __shared__ __align__(128) unsigned char oz_smem[];
    int8_t* sX = reinterpret_cast<int8_t*>(oz_smem);                 // [S][128*128]
    int8_t* sW = sX + (size_t)S * OZ_XTILE;                          // [S][64*128]
    __shared__ __align__(8) uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row_tile = blockIdx.x;
    const int nkc = p.K / OZ_KC;
    const int ct_begin = blockIdx.y * p.col_tiles_per_cta;
    const int ct_end = min(p.Nout / OZ_BN, ct_begin + p.col_tiles_per_cta);

    if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); mbar_fence_init(); }
    if (warp == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(dst));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor: D = s32, A = B = signed int8, both K-major, N = 64, M = 128
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);

    const int row = row_tile * OZ_BM + tid;                          // this thread's output row (TMEM lane tid)
    const bool row_ok = row < p.R;
    const double rs = row_ok ? p.rowscale[row] : 0.0;
    long long hrow = 0; int npts = 0;
    if (p.epi == EPI_QKV && row_ok) {
        if (row < p.rows0) { const int b = row / p.n0; npts = p.n0; hrow = (long long)b * HEADS * p.n0 + (row - b * p.n0); }
        else { const int r1 = row - p.rows0; const int b = r1 / p.n1; npts = p.n1;
               hrow = (long long)p.rows0 * HEADS + (long long)b * HEADS * p.n1 + (r1 - b * p.n1); }
    }

    unsigned load_phase = 0, mma_phase = 0;
    int x_resident_chunk = -1;
    for (int ct = ct_begin; ct < ct_end; ++ct) {
        for (int kc = 0; kc < nkc; ++kc) {
            if (tid == 0) {
                unsigned bytes = S * OZ_WTILE;
                const bool need_x = (x_resident_chunk != kc);
                if (need_x) bytes += S * OZ_XTILE;
                mbar_expect_tx(&bar_load, bytes);
                if (need_x) {
                    const int8_t* xsrc = p.Xs + ((size_t)(row_tile * nkc + kc) * S) * OZ_XTILE;
#pragma unroll
                    for (int s = 0; s < S; ++s) bulk_g2s(sX + (size_t)s * OZ_XTILE, xsrc + (size_t)s * OZ_XTILE, OZ_XTILE, &bar_load);
                }
                const int8_t* wsrc = p.Ws + ((size_t)(ct * nkc + kc) * S) * OZ_WTILE;
#pragma unroll
                for (int s = 0; s < S; ++s) bulk_g2s(sW + (size_t)s * OZ_WTILE, wsrc + (size_t)s * OZ_WTILE, OZ_WTILE, &bar_load);
                mbar_wait(&bar_load, load_phase);
                asm volatile("tcgen05.fence::after_thread_sync;");
                // all slice products of this k chunk, diagonal by diagonal
#pragma unroll
                for (int dd = 0; dd < S; ++dd) {
#pragma unroll
                    for (int s = 0; s <= dd; ++s) {
                        const int t = dd - s;
#pragma unroll
                        for (int kk = 0; kk < OZ_KC / 32; ++kk) {
                            const uint64_t da = umma_desc(sX + (size_t)s * OZ_XTILE + kk * 256, 128, OZ_KC * 8);
                            const uint64_t db = umma_desc(sW + (size_t)t * OZ_WTILE + kk * 256, 128, OZ_KC * 8);
                            umma_i8(tmem + dd * OZ_BN, da, db, idesc, (kc > 0 || s > 0 || kk > 0) ? 1u : 0u);
                        }
                    }
                }
                umma_commit(&bar_mma);          // arrives when every MMA issued so far has finished reading smem / writing TMEM
            }
            x_resident_chunk = (nkc == 1) ? 0 : kc;      // with one k chunk the X slices stay for all column tiles
            load_phase ^= 1;
            // everybody waits for the MMAs of this chunk before smem is overwritten / TMEM is read
            mbar_wait(&bar_mma, mma_phase);
            mma_phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;");
        }
        // ---- epilogue of column tile ct: Horner over the diagonals, 32 columns at a time
        const int n0 = ct * OZ_BN;
#pragma unroll 1
        for (int c = 0; c < OZ_BN / 32; ++c) {
            double t[32];
            uint32_t r[32];
            const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16) + c * 32;
            tmem_ld32(lane_addr + (S - 1) * OZ_BN, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = (double)(int)r[j];
#pragma unroll
            for (int dd = S - 2; dd >= 0; --dd) {
                tmem_ld32(lane_addr + dd * OZ_BN, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) t[j] = fma(t[j], 0.0078125, (double)(int)r[j]);
            }
            if (row_ok) {
                const int col0 = n0 + c * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    double y = t[j] * rs * p.colscale[col0 + j];
                    if (p.bias) y += p.bias[col0 + j];
                    if (p.relu) y = __double2hiint(y) < 0 ? 0.0 : y;
                    t[j] = y;
                }
                double* dst;
                if (p.epi == EPI_PLAIN) {
                    if (p.Res) {
                        const double* rr = p.Res + (long long)row * p.ldres + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 2) { const double2 v = *reinterpret_cast<const double2*>(rr + j); t[j] += v.x; t[j + 1] += v.y; }
                    }
                    dst = p.Y + (long long)row * p.ldy + col0;
                } else {
                    // 32 consecutive output channels = one head of q, k or v (head-major c' = h*32 + d)
                    const int which = col0 >> 7, h = (col0 & 127) >> 5;
                    double* base = which == 0 ? p.Qh : (which == 1 ? p.Kh : p.Vh);
                    dst = base + (hrow + (long long)h * npts) * (which == 2 ? LDH_V : LDH_QK);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(t[j], t[j + 1]);
            }
        }
        // TMEM is reused by the next column tile: all reads must be done before its first MMA
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

size_t ozaki_slices_bytes(int R, int K, int S) {
    const size_t rt = (size_t)((R + OZ_BM - 1) / OZ_BM);
    return rt * (size_t)(K / OZ_KC) * S * OZ_XTILE;
}

template <int S>
static cudaError_t slice_rows_t(const double* A0, int ld0, int K0, const double* A1, int ld1, int K1, int R,
                                int8_t* Xs, double* rowscale, cudaStream_t st) {
    const int K = K0 + K1;
    const long long Rpad = (long long)((R + OZ_BM - 1) / OZ_BM) * OZ_BM;
    const long long threads = Rpad * (K / 16);
    slice_rows_kernel<S><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(A0, ld0, K0, A1, ld1, K1, R, Xs, rowscale);
    return cudaGetLastError();
}

cudaError_t launch_slice_rows(const double* A0, int ld0, int K0, const double* A1, int ld1, int K1, int R, int S,
                              int8_t* Xs, double* rowscale, cudaStream_t st) {
    if (R <= 0) return cudaSuccess;
    const int K = K0 + K1;
    if ((K0 % OZ_KC) != 0 || (K1 % OZ_KC) != 0 || (K != 128 && K != 256 && K != 512)) return cudaErrorInvalidValue;
    cudaError_t e;
    switch (S) {
        case 6: e = slice_rows_t<6>(A0, ld0, K0, A1, ld1, K1, R, Xs, rowscale, st); break;
        case 7: e = slice_rows_t<7>(A0, ld0, K0, A1, ld1, K1, R, Xs, rowscale, st); break;
        case 8: e = slice_rows_t<8>(A0, ld0, K0, A1, ld1, K1, R, Xs, rowscale, st); break;
        default: return cudaErrorInvalidValue;
    }
    if (e == cudaSuccess) count_launch();
    return e;
}

template <int S>
static cudaError_t ozaki_gemm_t(const OzParams& p, dim3 grid, cudaStream_t st) {
    const size_t smem = (size_t)S * (OZ_XTILE + OZ_WTILE);
    cudaError_t e = cudaFuncSetAttribute(ozaki_gemm_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ozaki_gemm_kernel<S><<<grid, OZ_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_ozaki_gemm(const OzGemmArgs& a, int S, cudaStream_t st) {
    if (a.R <= 0) return cudaSuccess;
    if ((a.K % OZ_KC) != 0 || (a.Nout % OZ_BN) != 0) return cudaErrorInvalidValue;
    OzParams p;
    p.Xs = a.Xs; p.rowscale = a.rowscale; p.Ws = a.Ws; p.colscale = a.colscale; p.bias = a.bias;
    p.Res = a.Res; p.ldres = a.ldres; p.Y = a.Y; p.ldy = a.ldy; p.R = a.R; p.Nout = a.Nout; p.K = a.K;
    p.relu = a.relu; p.epi = a.epi; p.Qh = a.Qh; p.Kh = a.Kh; p.Vh = a.Vh; p.rows0 = a.rows0; p.n0 = a.n0; p.n1 = a.n1;
    const int row_tiles = (a.R + OZ_BM - 1) / OZ_BM, col_tiles = a.Nout / OZ_BN;
    // with one k chunk the X slices stay resident across column tiles: keep all column tiles in one CTA unless
    // that leaves SMs idle; otherwise every column tile reloads X anyway, so spread them out
    int groups = 1;
    if (a.K > OZ_KC) groups = col_tiles;
    else while (row_tiles * groups < 148 && groups < col_tiles) ++groups;
    p.col_tiles_per_cta = (col_tiles + groups - 1) / groups;
    dim3 grid(row_tiles, (col_tiles + p.col_tiles_per_cta - 1) / p.col_tiles_per_cta);
    cudaError_t e;
    switch (S) {
        case 6: e = ozaki_gemm_t<6>(p, grid, st); break;
        case 7: e = ozaki_gemm_t<7>(p, grid, st); break;
        case 8: e = ozaki_gemm_t<8>(p, grid, st); break;
        default: return cudaErrorInvalidValue;
    }
    if (e == cudaSuccess) count_launch();
    return e;
}

}  // namespace mdgat
