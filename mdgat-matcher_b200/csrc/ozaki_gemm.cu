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
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int OZ_BM = 128, OZ_BN = 32, OZ_KC = 128;
constexpr int OZ_XTILE = OZ_BM * OZ_KC, OZ_WTILE = OZ_BN * OZ_KC;       // bytes per slice tile
constexpr int OZ_WSTAGES = 3;                                             // W tile ring
constexpr int OZ_EPI_WARPS = 16, OZ_EPI_THREADS = OZ_EPI_WARPS * 32, OZ_THREADS = OZ_EPI_THREADS + 64;   // epilogue warps + MMA warp + loader warp
constexpr int OZ_MAXN = 384;                                               // widest GEMM (q/k/v stack): scales staged in smem

DEVINL int oz_canon(int r, int k) { return (r >> 3) * (OZ_KC * 8) + (k >> 4) * 128 + (r & 7) * 16 + (k & 15); }

DEVINL double pow2d(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }   // -1022 <= e <= 1023

// ---------------------------------------------------------------------------------------------------
// Slicing: one thread = 16 consecutive k of one row; the 8 lanes of a (row, 128-column chunk) share the
// chunk's exponent. Input = concat of up to two row-major float64 buffers (K0 + K1 columns, multiples
// of 128). rowscale[kchunk][Rpad] = 2^(e - 12).
// ---------------------------------------------------------------------------------------------------
// x: 16 consecutive k (group kt) of padded row r; the 8 lanes holding a (row, 128-column chunk) must be an aligned
// group of 8 lanes of a fully active warp
template <int S> DEVINL void cut_digits16(const double (&x)[16], int e, int8_t* base);
// MDGAT_SLICE_ONEFMA (read once; default 1): digits by cut_digits16() -- one FMA per value -- or by the telescoped roundings
// below (one or two FMAs per digit). Every slicer of a process uses the same rule, so the fused and stand-alone slicers agree.
static bool slice_onefma() {
    static const bool v = [] { const char* e = getenv("MDGAT_SLICE_ONEFMA"); return !(e && e[0] == '0'); }();
    return v;
}
template <int S>
DEVINL void slice_group(double (&x)[16], long long r, int kt, int Rpad, int8_t* Xs, double* rowscale, size_t chunk_stride, bool onefma) {
    const int k0 = kt * 16;
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) mx = fmax(mx, fabs(x[i]));
    // maximum over the 8 lanes that share this (row, 128-column chunk)
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);                         // mx = m * 2^e, m in [0.5, 1)  ->  |x| < 2^e
    e = max(-900, min(900, e));
    const int tile = (int)(r / OZ_BM), rr = (int)(r % OZ_BM);
    const int kchunk = k0 / OZ_KC, kk = k0 % OZ_KC;
    if ((kt & 7) == 0) rowscale[(size_t)kchunk * Rpad + r] = pow2d(e - 12);
    // Digits. With t = x * 2^(6-e) the most-significant-first cascade d_1 = rint(t), d_2 = rint((t - d_1) * 128), ..
    // telescopes: d_s = q_s - 128 q_(s-1) with q_s = rint(t * 128^(s-1)) (every step of the cascade is exact). Each q_s
    // is ONE FMA -- x * (2^(6-e) 128^(s-1)) + 1.5 * 2^52 leaves rint(.) in the low mantissa word (|q_s| < 2^48; only its
    // low 32 bits are needed because |d_s| <= 64) -- and the difference runs on the integer pipe: 7 FP64 instructions per
    // element instead of 28 (the slicer was FP64-bound), rint() / F2I never touch the quarter-rate conversion pipe.
    int8_t* base = Xs + (size_t)kchunk * chunk_stride + ((size_t)tile * S) * OZ_XTILE + oz_canon(rr, kk);
    if (onefma) { cut_digits16<S>(x, e, base); return; }
    // q_(s-1) is recomputed rather than kept (one more FMA per digit; 16 registers less: four CTAs per SM instead of
    // three -- the slicer is bound by the loads it keeps in flight, not by FP64 work)
#pragma unroll 1
    for (int s = 0; s < S; ++s) {
        const double cs = pow2d(6 - e + 7 * s);              // 6 - e + 7 s <= 6 + 900 + 42 < 1023
        const double cp = pow2d(6 - e + 7 * (s > 0 ? s - 1 : 0));
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int q = __double2loint(fma(x[i], cs, 6755399441055744.0));
            const int qp = s > 0 ? __double2loint(fma(x[i], cp, 6755399441055744.0)) : 0;
            const int d = q - (qp << 7);
            w[i >> 2] |= ((uint32_t)d & 0xffu) << (8 * (i & 3));
        }
        *reinterpret_cast<uint4*>(base + (size_t)s * OZ_XTILE) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---------------------------------------------------------------------------------------------------
// All S digits of 16 values from ONE FMA each (the forward path's digit cutter).
// I = rint(x 2^(6-e) 128^(S-1)), |I| <= 64 * 128^(S-1) because |x| < 2^e. The addend of the FMA is 1.5 * 2^52 + B with
// B = 64 * sum_k 128^k, so the low mantissa bits hold I + B >= 0, whose base-128 digits u_k in [0, 127] (the leading one in
// [0, 128]) are balanced digits shifted by 64: d_k = u_k - 64 in [-64, 63] (leading: [-64, 64]), sum_k d_k 128^k = I exactly --
// the same integer the telescoped roundings of slice_group() represent (x = 2^e sum_s d_s 2^(1-7s), |d_s| <= 64), in another
// valid digit set. Per plane the 7-bit fields of four values are moved into the bytes of one word (shift + PRMT) and 64 is
// subtracted from the four bytes at once (carry-free byte add of 0xC0): 1 FP64 + ~2.5 integer instructions per digit where the
// telescoped form spends 2 FP64 + 3 integer -- the slicer was issue-bound (ncu: 53 instructions per element, issue slots 60 %).
// ---------------------------------------------------------------------------------------------------
template <int S>
DEVINL void cut_digits16(const double (&x)[16], int e, int8_t* base) {
    static_assert(S >= 1 && S <= 7, "7 S + 1 bits must fit below the 2^51 bit of the rounding constant");
    constexpr unsigned long long B = 64ull * (((1ull << (7 * S)) - 1ull) / 127ull);
    const double cs = pow2d(6 - e + 7 * (S - 1));           // 6 - e + 7 (S - 1) <= 6 + 900 + 42 < 1023
    const double magic = 6755399441055744.0 + (double)B;    // an integer below 2^53: exact
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const double v = fma(x[i], cs, magic);
        lo[i] = (uint32_t)__double2loint(v);
        hi[i] = (uint32_t)__double2hiint(v);
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int sh = 7 * (S - 1 - s);                      // plane 0 = most significant digit
        uint32_t w[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t l = lo[4 * g + j], h = hi[4 * g + j];
                a[j] = sh == 0 ? l : (sh + 8 <= 32 ? (l >> sh) : (sh < 32 ? __funnelshift_r(l, h, sh) : (h >> (sh - 32))));
            }
            const uint32_t t01 = __byte_perm(a[0], a[1], 0x0040), t23 = __byte_perm(a[2], a[3], 0x0040);
            const uint32_t u = __byte_perm(t01, t23, 0x5410);                       // byte j = bits sh .. sh + 7 of value 4 g + j
            const uint32_t t = (u & 0x7f7f7f7fu) + 0x40404040u;                     // + 0xC0 per byte without carries between bytes
            // leading digit: all eight bits count (u_k up to 128); the others: bit 7 belongs to the next digit
            w[g] = s == 0 ? (t ^ (~u & 0x80808080u)) : (t ^ 0x80808080u);
        }
        *reinterpret_cast<uint4*>(base + (size_t)s * OZ_XTILE) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <int S>
__global__ void __launch_bounds__(128, 8)
slice_rows_kernel(const double* __restrict__ A0, int ld0, int K0, const double* __restrict__ A1, int ld1, int K1,
                  int R, int8_t* __restrict__ Xs, double* __restrict__ rowscale, size_t chunk_stride, bool onefma) {
    const int K = K0 + K1, tpr = K / 16;                 // threads per row
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = gid / tpr;
    const int kt = (int)(gid - r * tpr);                 // which 16-wide k group
    const int Rpad = ((R + OZ_BM - 1) / OZ_BM) * OZ_BM;
    if (r >= Rpad) return;                               // whole warps: Rpad * tpr is a multiple of 1024
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
    slice_group<S>(x, r, kt, Rpad, Xs, rowscale, chunk_stride, onefma);
}

// ---------------------------------------------------------------------------------------------------
// The forward path's slicer. slice_rows_kernel above gives every thread 128 contiguous bytes of a row, so each of its
// load instructions touches 32 different cache lines and each plane store half-fills eight: the kernel sat at 65-80 % of
// the L1 wavefront pipe and a third of the HBM rate (ncu, profiles/r2_*). Here a CTA owns 64 rows x one 128-column chunk:
// the 64 row segments (1 KB each, contiguous) arrive by TMA bulk copies into shared memory with a 1040-byte row pitch
// (16 bytes past a multiple of 128: the 8 x LDS.128 of a warp whose lanes are 32 consecutive rows are conflict-free),
// thread (row, kt) cuts 16 consecutive k, the chunk maximum of a row is combined across the 8 warps that share it through
// shared memory, and a warp's 16-byte plane stores fall on 4 full 128-byte lines (8 rows x 16 bytes are contiguous in the
// core-matrix layout). Same digits, same layout, bit for bit.
// ---------------------------------------------------------------------------------------------------
constexpr int SL_ROWS = 32, SL_PITCH = 1040, SL_THREADS = 256, SL_STAGES = 2;
// Persistent: the CTAs (three per SM) walk the (row tile, k chunk) tiles round robin through a two-stage ring -- the rows of
// the next tile travel while the current one is cut -- so neither the HBM latency of a tile nor the last partial wave of a
// one-tile-per-CTA grid (1.73 waves at cfg2) is exposed.
template <int S, bool ONEFMA>
__global__ void __launch_bounds__(SL_THREADS, 3)
slice_rows_tiled_kernel(const double* __restrict__ A0, int ld0, int K0, const double* __restrict__ A1, int ld1,
                        int R, int nchunks, int8_t* __restrict__ Xs, double* __restrict__ rowscale, size_t chunk_stride) {
    extern __shared__ __align__(128) unsigned char sl_smem[];                        // [stage][32 rows][1040 B] | s_mx[8][32]
    __shared__ __align__(8) uint64_t full[SL_STAGES];
    double* s_mx = reinterpret_cast<double*>(sl_smem + SL_STAGES * SL_ROWS * SL_PITCH);
    const int tid = threadIdx.x, rl = tid & 31, kt = tid >> 5;
    const int Rpad = ((R + OZ_BM - 1) / OZ_BM) * OZ_BM;
    const int ntiles = (Rpad / SL_ROWS) * nchunks;
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    // tile t = (row tile t / nchunks, k chunk t % nchunks); warp 0 issues the row copies of a tile (one per lane)
    auto issue = [&](int t, int stage) {
        const long long row0 = (long long)(t / nchunks) * SL_ROWS;
        const int k0 = (t % nchunks) * OZ_KC;
        const int nvalid = (int)max(0ll, min((long long)SL_ROWS, (long long)R - row0));
        if (rl == 0) mbar_expect_tx(&full[stage], (unsigned)nvalid * 1024u);       // nvalid == 0: a plain arrival
        __syncwarp();
        if (rl < nvalid) {
            const double* src = k0 < K0 ? A0 + (row0 + rl) * ld0 + k0 : A1 + (row0 + rl) * ld1 + (k0 - K0);
            bulk_g2s(sl_smem + (stage * SL_ROWS + rl) * SL_PITCH, src, 1024, &full[stage]);
        }
    };
    int t = blockIdx.x;
    if (kt == 0 && t < ntiles) issue(t, 0);
    for (int it = 0; t < ntiles; t += gridDim.x, ++it) {
        const int stage = it & 1;
        if (kt == 0 && t + (int)gridDim.x < ntiles) issue(t + gridDim.x, stage ^ 1);   // that buffer was released by the barrier below
        const long long row0 = (long long)(t / nchunks) * SL_ROWS, r = row0 + rl;
        const int kchunk = t % nchunks;
        const int nvalid = (int)max(0ll, min((long long)SL_ROWS, (long long)R - row0));
        mbar_wait(&full[stage], (unsigned)((it >> 1) & 1));
        double x[16];
        if (rl < nvalid) {
            const double* srow = reinterpret_cast<const double*>(sl_smem + (stage * SL_ROWS + rl) * SL_PITCH + kt * 128);
#pragma unroll
            for (int i = 0; i < 16; i += 2) { const double2 v = *reinterpret_cast<const double2*>(srow + i); x[i] = v.x; x[i + 1] = v.y; }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = 0.0;
        }
        double mx = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmax(mx, fabs(x[i]));
        s_mx[kt * SL_ROWS + rl] = mx;
        __syncthreads();                                     // chunk maxima visible; everybody has read its part of the stage
#pragma unroll
        for (int q = 0; q < 8; ++q) mx = fmax(mx, s_mx[q * SL_ROWS + rl]);
        int e = 0;
        if (mx > 0.0) frexp(mx, &e);                         // mx = m * 2^e, m in [0.5, 1)  ->  |x| < 2^e
        e = max(-900, min(900, e));
        if (kt == 0) rowscale[(size_t)kchunk * Rpad + r] = pow2d(e - 12);
        const int tile = (int)(r / OZ_BM), rr = (int)(r % OZ_BM);
        int8_t* base = Xs + (size_t)kchunk * chunk_stride + ((size_t)tile * S) * OZ_XTILE + oz_canon(rr, kt * 16);
        if constexpr (ONEFMA) cut_digits16<S>(x, e, base);
        else
        // digits as in slice_group(): d_s = q_s - 128 q_(s-1), q_s = rint(x 2^(6-e) 128^s) read out of the mantissa
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
            const double cs = pow2d(6 - e + 7 * s);
            const double cp = pow2d(6 - e + 7 * (s > 0 ? s - 1 : 0));
            uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int q = __double2loint(fma(x[i], cs, 6755399441055744.0));
                const int qp = s > 0 ? __double2loint(fma(x[i], cp, 6755399441055744.0)) : 0;
                const int d = q - (qp << 7);
                w[i >> 2] |= ((uint32_t)d & 0xffu) << (8 * (i & 3));
            }
            *reinterpret_cast<uint4*>(base + (size_t)s * OZ_XTILE) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncthreads();                                     // s_mx is rewritten by the next tile
    }
}

// ---------------------------------------------------------------------------------------------------
// tcgen05 helpers (raw PTX)
// ---------------------------------------------------------------------------------------------------
DEVINL uint64_t umma_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
template <bool ACCUMULATE>
DEVINL void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
    if (ACCUMULATE)
        asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\n"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                     :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                     :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
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
    const int8_t* Xs[2]; const double* rowscale[2];    // per 128-column k chunk: slice planes [row_tile][S][128*128] + 2^(e - 12) per row
    const int8_t* Ws; const double* colscale;          // sliced weights + 2^(f) per (k chunk, column)
    const double* bias;                                // [Nout] or null
    const double* Res; int ldres;                      // residual (may alias Y) or null
    double* Y; int ldy;
    int R, Nout, K;                                    // K multiple of 128, Nout multiple of 32
    int relu;
    int epi;                                           // EPI_PLAIN / EPI_QKV
    double *Qh, *Kh, *Vh; int rows0, n0, n1;
    int col_tiles_per_cta;                             // blockIdx.y selects a group of column tiles
    int items_per_cta;                                 // > 0: persistent mode, CTA c walks the (row tile, column tile) items
                                                       // [c * items_per_cta, (c + 1) * items_per_cta) in row-major order
    // fused slicing of the output (EPI_PLAIN, CTA owns whole rows of Y): the digit planes of this CTA's 128 x Nout tile of Y,
    // i.e. the operand of the next GEMM, in the layout launch_slice_rows() writes; null = off
    int8_t* slice_out; double* slice_scale; unsigned long long slice_chunk_stride; int slice_onefma;
    long long* trace;                                  // debug timeline (null = off)
    int dbg;                                           // debug switches (mdgat_debug_flags)
};

// One CTA = one 128-row tile x a group of 32-column tiles (all of them when there are >= 148 row tiles), 18 warps:
//   warp 17, lane 0  loader. For every k chunk one TMA bulk copy brings in the S slice planes of X (they stay for all
//                    column tiles of the chunk); W planes arrive through a 3-deep ring of bulk copies.
//   warp 16, lane 0  MMA issuer: per unit (k chunk, column tile) the S(S+1)/2 slice products as S stacked-N instructions
//                    per k step into one of TWO TMEM accumulator sets (S diagonals x 32 columns each); tcgen05.commit
//                    hands the set to the epilogue and the W stage back to the ring.
//   warps 0..15      epilogue (see the comment there): TMEM -> int32 merges -> float64 Horner -> scales; k chunks are
//                    accumulated in float64 through the output buffer, each with its own row/column scale; bias, ReLU,
//                    residual or the q/k/v scatter after the last chunk. While one accumulator set is drained the
//                    tensor core fills the other.
// Template switches: EPI (plain / q,k,v scatter), DBG (0 production, 1 timeline probes in the loader + MMA threads,
// 2 probes and debug switches in the epilogue too), FULL (row count a multiple of 128: no row checks), GROUPS (1: all
// 16 epilogue warps on every unit, 2: warps 0-7 on even units / set 0, warps 8-15 on odd units / set 1).
template <int S, int EPI, int DBG, bool FULL, int GROUPS>
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm_kernel(const __grid_constant__ OzParams p) {
    extern __shared__ __align__(128) unsigned char oz_smem[];
    int8_t* sX = reinterpret_cast<int8_t*>(oz_smem);                 // [S][128*128]
    int8_t* sW = sX + (size_t)S * OZ_XTILE;                          // [OZ_WSTAGES][S][32*128]
    // x_full / x_free are per slice plane: the first unit of a k chunk starts on plane 0 while planes 1.. are still in
    // flight, and the last unit hands the planes back one by one so that the next chunk's reload overlaps its MMAs
    __shared__ __align__(8) uint64_t x_full[S], x_free[S], w_full[OZ_WSTAGES], w_empty[OZ_WSTAGES], tm_full[2], tm_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) double s_cs[2 * OZ_MAXN], s_bias[OZ_MAXN];             // column scales per k chunk, bias
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nkc = p.K / OZ_KC;
    const int Rpad = ((p.R + OZ_BM - 1) / OZ_BM) * OZ_BM;
    // Work of this CTA: a range of (row tile, column tile) items in row-major order, cut into segments that stay inside
    // one row tile. Classic mode: one segment (row tile blockIdx.x, column tiles of group blockIdx.y). Persistent mode
    // (one CTA per SM, items_per_cta > 0): up to three segments, which balances the SMs when the row tiles are not a
    // multiple of the SM count (256 row tiles on 148 SMs: 2 waves of which the second is 73 % full).
    const int nct_all = p.Nout / OZ_BN;
    int item_begin, item_end;
    if (p.items_per_cta > 0) {
        item_begin = blockIdx.x * p.items_per_cta;
        item_end = min(((p.R + OZ_BM - 1) / OZ_BM) * nct_all, item_begin + p.items_per_cta);
    } else {
        const int cb = blockIdx.y * p.col_tiles_per_cta;
        item_begin = blockIdx.x * nct_all + cb;
        item_end = blockIdx.x * nct_all + min(nct_all, cb + p.col_tiles_per_cta);
    }

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < S; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_free[i], 1); }
#pragma unroll
        for (int i = 0; i < OZ_WSTAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
#pragma unroll
        for (int i = 0; i < 2; ++i) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], OZ_EPI_THREADS / GROUPS); }
        mbar_fence_init();
    }
    for (int i = tid; i < nkc * p.Nout; i += OZ_THREADS) s_cs[i] = p.colscale[i];
    for (int i = tid; i < p.Nout; i += OZ_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.0;
    if (warp == OZ_EPI_WARPS) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(dst));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    pdl_wait();                                                      // up to here only weights (scales, bias) were read
    pdl_trigger();
    constexpr int TM_SET = S * OZ_BN;                                // TMEM columns of one accumulator set
    Tracer tr;
    if (DBG) tr.init(p.trace, warp == OZ_EPI_WARPS + 1 ? 0 : (warp == OZ_EPI_WARPS ? 1 : (warp == 0 ? 2 : 3)), blockIdx.x == 0 && blockIdx.y == 0 && (tid & 31) == 0 && (warp >= OZ_EPI_WARPS || (DBG > 1 && (warp == 0 || warp == 4))));

    if (warp == OZ_EPI_WARPS + 1) {
        // ------------------------------------------------------------------ loader: TMA bulk copies, runs ahead
        if (elect_one()) {
            // (kc, ctl, stage, ring phase) are walked incrementally: a division by the runtime nct costs this single
            // thread ~100 cycles of dependent instructions per unit. ug / xl count units and X chunk loads over all
            // segments of the CTA (barrier phases run on).
            int stage = 0, ug = 0, xl = 0; unsigned wphase = 1;       // wphase: parity of the ring pass before this one
            for (int it = item_begin; it < item_end;) {
                const int row_tile = it / nct_all, ct_begin = it - row_tile * nct_all;
                const int nct = min(nct_all - ct_begin, item_end - it), units = nkc * nct;
                it += nct;
                int kc = 0, ctl = 0;
                for (int u = 0; u < units; ++u, ++ug) {
                    const int ct = ct_begin + ctl;
                    if (DBG && !(p.dbg & 256)) tr.mark(1000 + ug);
                    // the W tile first: the X planes below may have to wait for the previous chunk's last MMAs
                    if (ug >= OZ_WSTAGES) mbar_wait(&w_empty[stage], wphase);
                    mbar_expect_tx(&w_full[stage], S * OZ_WTILE);
                    const int8_t* wsrc = p.Ws + ((size_t)(ct * nkc + kc) * S) * OZ_WTILE;
                    bulk_g2s(sW + (size_t)stage * S * OZ_WTILE, wsrc, S * OZ_WTILE, &w_full[stage]);
                    if (ctl == 0) {
                        const int8_t* xsrc = (kc == 0 ? p.Xs[0] : p.Xs[1]) + ((size_t)row_tile * S) * OZ_XTILE;
#pragma unroll
                        for (int sp = 0; sp < S; ++sp) {
                            if (xl > 0) mbar_wait(&x_free[sp], (unsigned)((xl - 1) & 1));   // last MMA reading this plane is done
                            mbar_expect_tx(&x_full[sp], OZ_XTILE);
                            bulk_g2s(sX + (size_t)sp * OZ_XTILE, xsrc + (size_t)sp * OZ_XTILE, OZ_XTILE, &x_full[sp]);
                        }
                        ++xl;
                    }
                    if (DBG && !(p.dbg & 512)) tr.mark(2000 + ug);
                    if (++ctl == nct) { ctl = 0; ++kc; }
                    if (++stage == OZ_WSTAGES) { stage = 0; wphase ^= 1u; }
                }
            }
        }
    } else if (warp == OZ_EPI_WARPS) {
        // ------------------------------------------------------------------ MMA issuer: one thread, never waits on loads it issued
        // (elected with elect.sync: with a `lane == 0` test the compiler wraps every tcgen05.mma in its own
        // ELECT / BRA.U.ANY retry loop -- 11 instructions between two MMAs instead of 4)
        if (elect_one()) {
            // Stacked-N issue. The S weight planes of a column tile sit back to back in shared memory, i.e. they form ONE
            // K-major operand of S*32 rows. Multiplying activation plane s by planes 0..S-1-s in a single MMA of
            // N = (S-s)*32 and writing it 32*s columns into the accumulator set drops product (s, t) onto diagonal
            // s + t: the same S(S+1)/2 slice products as before, but in S wide instructions per k step instead of
            // S(S+1)/2 narrow ones. A 128x32x32 MMA takes 45 cycles on this part (operand fetch bound, 16 in
            // theory), a 128xNx32 one N/2 cycles from N = 128 up (tools/ubench/umma_i8_rate.cu).
            const uint64_t xd0 = umma_desc(sX, 128, OZ_KC * 8), wd0 = umma_desc(sW, 128, OZ_KC * 8);
            int stage = 0, ug = 0, xl = 0; unsigned wphase = 0;
            for (int it = item_begin; it < item_end;) {
              const int ct_begin = it % nct_all;
              const int nct = min(nct_all - ct_begin, item_end - it), units = nkc * nct;
              it += nct;
              int ctl = 0;
              for (int uu = 0; uu < units; ++uu, ++ug) {
                const int u = ug;                                    // global unit index of this CTA
                const int set = u & 1;
                if (DBG && !(p.dbg & 1024)) tr.mark(3000 + u);
                const bool first_of_chunk = ctl == 0, last_of_chunk = ctl == nct - 1;
                mbar_wait(&w_full[stage], wphase);
                if (DBG && !(p.dbg & 2048)) tr.mark(4000 + u);
                if (u >= 2) mbar_wait(&tm_empty[set], (unsigned)((u / 2 - 1) & 1));   // epilogue drained this accumulator set
                if (DBG && !(p.dbg & 4096)) tr.mark(5000 + u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint64_t wdu = wd0 + (uint64_t)((stage * S * OZ_WTILE) >> 4);
                const uint32_t dbase = tmem + set * TM_SET;
                // D = s32, A = B = signed int8, both K-major, M = 128, N = (S - s) * 32
                auto issue = [&](int s, int kk) {
                    constexpr uint32_t ibase = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
                    const uint32_t idesc = ibase | ((uint32_t)(((S - s) * OZ_BN) >> 3) << 17);
                    const uint64_t da = xd0 + (uint64_t)((s * OZ_XTILE + kk * 256) >> 4);
                    const uint64_t db = wdu + (uint64_t)((kk * 256) >> 4);
                    if (s > 0 || kk > 0) umma_i8<true>(dbase + s * OZ_BN, da, db, idesc);
                    else umma_i8<false>(dbase, da, db, idesc);
                };
                if (first_of_chunk || last_of_chunk) {
                    // plane-major order: plane s is needed only once planes < s are done (first unit of a chunk: the
                    // planes arrive one after the other) and is handed back as soon as its four k steps are issued
                    // (last unit: the loader refills it while the remaining planes are multiplied)
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (first_of_chunk) mbar_wait(&x_full[s], (unsigned)(xl & 1));
#pragma unroll
                        for (int kk = 0; kk < OZ_KC / 32; ++kk) issue(s, kk);
                        if (last_of_chunk) umma_commit(&x_free[s]);
                    }
                    if (first_of_chunk) ++xl;
                } else {
#pragma unroll
                    for (int kk = 0; kk < OZ_KC / 32; ++kk)
#pragma unroll
                        for (int s = 0; s < S; ++s) issue(s, kk);
                }
                umma_commit(&tm_full[set]);                 // accumulator set ready for the epilogue
                umma_commit(&w_empty[stage]);               // W stage free once these MMAs have read it
                if (DBG && !(p.dbg & 8192)) tr.mark(6000 + u);
                if (++ctl == nct) ctl = 0;
                if (++stage == OZ_WSTAGES) { stage = 0; wphase ^= 1u; }
              }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        // 16 warps = 4 per SM sub-partition. Warp w owns TMEM lanes 32*(w%4).. (hardware rule) and columns 8*(w/4)..+7
        // of every 32-column unit. Accumulators are read with the 16x256b shape, i.e. in MMA C-fragment order: lane t
        // holds rows t/4 and t/4 + 8 of a 16-lane half and the column pair 2*(t%4), +1 -- so a quad owns 64
        // contiguous bytes of an output row and a warp-wide double2 access touches 8 rows x 64 B (19.7 B/clk/SM of
        // store throughput against 9.1 for the thread-per-row shape, tools/ubench/store_patterns.cu).
        // Register budget: 18 warps cap a thread at 96 registers and every spill is a local-memory load queued behind
        // the global stores, so the diagonals are fetched in two rounds and neighbouring ones merged in int32 at once.
        // GROUPS == 2: warps 0-7 take the even units (accumulator set 0), warps 8-15 the odd ones (set 1), each warp
        // covering 16 columns in two passes of 8 -- the two groups are then in different phases (TMEM latency, FP64
        // math, stores) at any time instead of all 16 warps queueing for the same pipe.
        constexpr int PASSES = GROUPS;
        const int lane = tid & 31, grp = GROUPS == 2 ? warp >> 3 : 0, wl = GROUPS == 2 ? warp & 7 : warp;
        const int quarter = wl & 3, cg0 = (wl >> 2) * PASSES;
        constexpr int G = (S + 1) / 2;                               // digit groups: S odd: {0}, {1,2}, {3,4}, ..; S even: {0,1}, {2,3}, ..
        const double MAGIC = 6755399441055744.0;                     // 1.5 * 2^52
        const bool res_first = p.Res && !p.relu;
        const bool one_chunk = nkc == 1;
        const int* s_cs_hi = reinterpret_cast<const int*>(s_cs) + 1; // high words of the column scales (exact powers of two)
        int ug_base = 0;                                             // units of the segments before this one (accumulator set / phase)
      for (int it = item_begin; it < item_end;) {
        const int row_tile = it / nct_all, ct_begin = it - row_tile * nct_all;
        const int nct = min(nct_all - ct_begin, item_end - it), units = nkc * nct;
        it += nct;
        const int rbase = row_tile * OZ_BM + quarter * 32 + (lane >> 2);   // rows rbase + 8 i, i = 0..3
        int hrow[EPI == EPI_QKV ? 4 : 1], npts[EPI == EPI_QKV ? 4 : 1];      // head-major row of the point, points per set
        if (EPI == EPI_QKV) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = rbase + 8 * i;
                hrow[i] = 0; npts[i] = 0;
                if (FULL || row < p.R) {
                    if (row < p.rows0) { const int b = row / p.n0; npts[i] = p.n0; hrow[i] = b * HEADS * p.n0 + (row - b * p.n0); }
                    else { const int r1 = row - p.rows0; const int b = r1 / p.n1; npts[i] = p.n1;
                           hrow[i] = p.rows0 * HEADS + b * HEADS * p.n1 + (r1 - b * p.n1); }
                }
            }
        }
        // element offsets fit 32 bits (checked by the launcher): one IMAD per address instead of 64-bit chains
        const int yrow = rbase * p.ldy, ystep = 8 * p.ldy, rrow = rbase * p.ldres, rstep = 8 * p.ldres;
        // this group's units of the segment: those whose CTA-wide index ug_base + us has the group's parity
        const int us0 = GROUPS == 2 ? ((grp - ug_base) & 1) : 0;
        int kc = 0, ctl = us0;                                       // unit (kc, ct_begin + ctl), walked incrementally
        for (int us = us0; us < units; us += GROUPS, ctl += GROUPS) {
            while (ctl >= nct) { ctl -= nct; ++kc; }
            const int u = ug_base + us;                              // CTA-wide unit index
            const int set = u & 1;
            const bool first = kc == 0, last = kc == nkc - 1;
            int rsh[4];                                              // rsh: high word of the row scale 2^(e-12)
            const int* rsp = reinterpret_cast<const int*>(kc == 0 ? p.rowscale[0] : p.rowscale[1]) + 1 + 2 * rbase;
#pragma unroll
            for (int i = 0; i < 4; ++i) rsh[i] = (FULL || rbase + 8 * i < p.R) ? rsp[16 * i] : 0;
#pragma unroll
          for (int pz = 0; pz < PASSES; ++pz) {
            const int cg = cg0 + pz;
            const int col0 = (ct_begin + ctl) * OZ_BN + cg * 8 + (lane & 3) * 2;
            // operands that come from global memory are requested before the wait on the tensor core
            double2 add[4];
            const double2 bias2 = *reinterpret_cast<const double2*>(s_bias + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                add[i] = one_chunk ? bias2 : make_double2(0.0, 0.0);
                if (FULL || rbase + 8 * i < p.R) {
                    if (EPI == EPI_PLAIN) {
                        if (!first) add[i] = *reinterpret_cast<const double2*>(p.Y + (yrow + i * ystep + col0));
                        else if (res_first) {
                            const double2 v = *reinterpret_cast<const double2*>(p.Res + (rrow + i * rstep + col0));
                            add[i].x += v.x; add[i].y += v.y;
                        }
                    }
                }
            }
            if (pz == 0) {
                if (DBG > 1) tr.mark(7000 + u);
                mbar_wait(&tm_full[set], (unsigned)((u / 2) & 1));
                if (DBG > 1) tr.mark(8000 + u);
                asm volatile("tcgen05.fence::after_thread_sync;");
            }
            // m[g][half * 4 + (row & 8 ? 2 : 0) + column]: group g of the diagonals merged exactly in int32
            // (|acc_dd| <= (dd+1) * 128 * 64 * 64 < 2^23, so acc_dd * 128 + acc_dd+1 < 2^31)
            int m[G][8];
            const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + set * TM_SET + cg * 8;
            auto ld_diag = [&](int dd, int half, uint32_t (&r)[4]) {
                asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr + ((uint32_t)(half * 16) << 16) + dd * OZ_BN));
            };
            auto fetch_groups = [&](int g_lo, int g_hi) {            // groups [g_lo, g_hi): one round of TMEM loads
                uint32_t raw[2][2][2][4];                            // [group][digit][half][4]
#pragma unroll
                for (int g = g_lo; g < g_hi; ++g) {
                    const int d0 = (S & 1) ? 2 * g - 1 : 2 * g;      // first diagonal of the group (-1: group 0 of an odd S is diagonal 0 alone)
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (d0 >= 0) ld_diag(d0, half, raw[g - g_lo][0][half]);
                        ld_diag(d0 + 1, half, raw[g - g_lo][1][half]);
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int g = g_lo; g < g_hi; ++g) {
                    const int d0 = (S & 1) ? 2 * g - 1 : 2 * g;
#pragma unroll
                    for (int half = 0; half < 2; ++half)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            m[g][half * 4 + j] = d0 >= 0 ? (int)raw[g - g_lo][0][half][j] * 128 + (int)raw[g - g_lo][1][half][j]
                                                         : (int)raw[g - g_lo][1][half][j];
                }
            };
            fetch_groups(G - 2 > 0 ? G - 2 : 0, G);                   // the two least significant groups
            if (G > 2) fetch_groups(0, G - 2);
            if (pz == PASSES - 1) {
                // all TMEM reads of this set are complete: hand it back to the tensor core
                asm volatile("tcgen05.fence::before_thread_sync;");
                mbar_arrive(&tm_empty[set]);
                if (DBG > 1) tr.mark(9000 + u);
            }
            // Horner over the groups in float64; int32 -> float64 with the 2^52 magic constant (integer ALU + one DADD
            // instead of a quarter-rate I2F.F64)
            auto to_f64 = [&](int v) {
                return __hiloint2double(0x43380000 + (v >> 31), v) - MAGIC;      // exact for any int32
            };
            double t[8];
            // S <= 6: the merged groups fit ONE 64-bit integer (S = 5: 23 + 28 = 51 bits), combined by integer multiply-adds and
            // converted once (I2F.F64.S64) -- one conversion-pipe instruction instead of G DADDs and G - 1 DFMAs on the FP64 pipe,
            // which the epilogue warps were queueing for (math-pipe throttle was their second largest stall). The value is
            // 2^(14 (G - 1)) times the Horner sum; the factor goes into the exponent assembly below.
            constexpr bool I64 = S <= 6;
            constexpr int HSHIFT = I64 ? 14 * (G - 1) : 0;
            if (DBG > 1 && (p.dbg & 2)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { int x = 0;
#pragma unroll
                    for (int g = 0; g < G; ++g) x ^= m[g][j];
                    t[j] = __hiloint2double(x, x); }
            } else if (I64) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    long long v = (long long)m[G - 1][j];
#pragma unroll
                    for (int g = G - 2; g >= 0; --g) {
                        if (g == G - 2) asm("mad.wide.s32 %0, %1, 16384, %0;" : "+l"(v) : "r"(m[g][j]));
                        else if (g == G - 3) asm("mad.wide.s32 %0, %1, 268435456, %0;" : "+l"(v) : "r"(m[g][j]));
                        else v += (long long)m[g][j] << (14 * (G - 1 - g));
                    }
                    t[j] = __ll2double_rn(v);
                }
            } else
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                double h = to_f64(m[G - 1][j]);
#pragma unroll
                for (int g = G - 2; g >= 0; --g) h = fma(h, 0.00006103515625, to_f64(m[g][j]));   // 2^-14 per group
                t[j] = h;                                             // even S: group 0 carries acc_0 * 128, folded into the scale below
            }
            // scale 2^(e_row - 12) * 2^(f_col) [* 2^-7] assembled in the exponent field: both factors are exact powers of two
            const int csh0 = s_cs_hi[2 * (kc * p.Nout + col0)] - (0x3FF00000 + ((((S & 1) ? 0 : 7) + HSHIFT) << 20));
            const int csh1 = s_cs_hi[2 * (kc * p.Nout + col0 + 1)] - (0x3FF00000 + ((((S & 1) ? 0 : 7) + HSHIFT) << 20));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!FULL && rbase + 8 * i >= p.R) continue;
                double2 y;
                y.x = fma(t[2 * i], __hiloint2double(rsh[i] + csh0, 0), add[i].x);
                y.y = fma(t[2 * i + 1], __hiloint2double(rsh[i] + csh1, 0), add[i].y);
                if (DBG > 1 && (p.dbg & 1) && y.x != 1234.5) continue;
                if (EPI == EPI_QKV) {
                    // single k chunk; 32 consecutive output channels = one head of q, k or v (head-major c' = h*32 + d)
                    const int which = col0 >> 7, h = (col0 & 127) >> 5;
                    double* base = which == 0 ? p.Qh : (which == 1 ? p.Kh : p.Vh);
                    *reinterpret_cast<double2*>(base + (long long)(hrow[i] + h * npts[i]) * (which == 2 ? LDH_V : LDH_QK) + (col0 & 31)) = y;
                    continue;
                }
                if (last) {
                    if (!one_chunk) { y.x += bias2.x; y.y += bias2.y; }
                    if (p.relu) {
                        y.x = __double2hiint(y.x) < 0 ? 0.0 : y.x;
                        y.y = __double2hiint(y.y) < 0 ? 0.0 : y.y;
                        if (p.Res) {        // launcher guarantees Res != Y here when there are several k chunks
                            const double2 v = *reinterpret_cast<const double2*>(p.Res + (rrow + i * rstep + col0));
                            y.x += v.x; y.y += v.y;
                        }
                    }
                }
                *reinterpret_cast<double2*>(p.Y + (yrow + i * ystep + col0)) = y;
            }
          }
        }
        ug_base += units;
      }
        if (EPI == EPI_PLAIN && p.slice_out != nullptr) {
            const int row_tile = blockIdx.x;                         // classic mode only (checked by the launcher)
            // Tail: the digit planes of this CTA's rows of Y for the next GEMM. The rows were written by other warps of
            // this CTA: the named barrier orders those stores before the loads below (same SM, coherent L1).
            asm volatile("bar.sync 1, %0;" :: "n"(OZ_EPI_THREADS) : "memory");
            const int tpr = p.Nout / 16;                                 // 16-wide groups per row
            const double* Yc = p.Y;                                      // plain loads: Y was written in this kernel
            for (int task = tid; task < OZ_BM * tpr; task += OZ_EPI_THREADS) {
                const int rr = task / tpr, kt = task - rr * tpr;
                const long long r = (long long)row_tile * OZ_BM + rr;
                double x[16];
                if (r < p.R) {
                    const double* src = Yc + r * p.ldy + kt * 16;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) { const double2 v = *reinterpret_cast<const double2*>(src + i); x[i] = v.x; x[i + 1] = v.y; }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] = 0.0;
                }
                slice_group<S>(x, r, kt, Rpad, p.slice_out, p.slice_scale, (size_t)p.slice_chunk_stride, p.slice_onefma != 0);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == OZ_EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

// ---------------------------------------------------------------------------------------------------
// Issue-rate ceiling of kind::i8 (roofline denominator of the tcgen05 engines, measured live by bench.py): one CTA
// per SM, one thread issues 128x256x32 MMAs back to back on operands resident in shared memory.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) i8_peak_kernel(int iters) {
    extern __shared__ __align__(128) unsigned char pk_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(pk_smem)[i] = 0x01010101u * (i & 3);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(dst));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da0 = umma_desc(pk_smem, 128, 1024), db0 = umma_desc(pk_smem + 16384, 128, 1024);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int kk = j & 3;
                umma_i8<true>(tmem, da0 + (uint64_t)((kk * 256) >> 4), db0 + (uint64_t)((kk * 256) >> 4), idesc);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

cudaError_t measure_i8_peak(double* tops) {
    int dev = 0, sms = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const size_t smem = (128 + 256) * 128;
    if ((e = cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    const int iters = 2000;
    float ms = 0.f;
    for (int rep = 0; rep < 2; ++rep) {         // first rep warms up
        cudaEventRecord(t0);
        i8_peak_kernel<<<sms, 128, smem>>>(iters);
        cudaEventRecord(t1);
        if ((e = cudaEventSynchronize(t1)) != cudaSuccess) break;
        cudaEventElapsedTime(&ms, t0, t1);
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    if (e != cudaSuccess) return e;
    *tops = (double)sms * iters * 16.0 * 2.0 * 128 * 256 * 32 / (ms * 1e-3) / 1e12;
    return cudaGetLastError();
}

// the GEMM can cut its own output into digit planes when one CTA owns whole rows of Y (no column split)
bool ozaki_gemm_can_slice(int R, int Nout) { return (R + OZ_BM - 1) / OZ_BM >= 148 && (Nout % OZ_KC) == 0; }

size_t ozaki_slices_bytes(int R, int K, int S) {
    const size_t rt = (size_t)((R + OZ_BM - 1) / OZ_BM);
    return rt * (size_t)(K / OZ_KC) * S * OZ_XTILE;
}

template <int S>
static cudaError_t slice_rows_t(const double* A0, int ld0, int K0, const double* A1, int ld1, int K1, int R,
                                int8_t* Xs, double* rowscale, size_t chunk_stride, cudaStream_t st) {
    const int K = K0 + K1;
    const long long Rpad = (long long)((R + OZ_BM - 1) / OZ_BM) * OZ_BM;
    // MDGAT_SLICE_TILED=0 (read once) falls back to the thread-per-128-bytes kernel; the tiled one needs 16-byte aligned rows
    static const bool tiled = [] { const char* e = getenv("MDGAT_SLICE_TILED"); return !(e && e[0] == '0'); }();
    const bool aligned = (ld0 % 2) == 0 && (reinterpret_cast<uintptr_t>(A0) % 16) == 0 &&
                         (A1 == nullptr || ((ld1 % 2) == 0 && (reinterpret_cast<uintptr_t>(A1) % 16) == 0));
    if (tiled && aligned) {
        const size_t smem = (size_t)SL_STAGES * SL_ROWS * SL_PITCH + 8 * SL_ROWS * sizeof(double);
        const bool one = slice_onefma();
        cudaError_t e = one ? cudaFuncSetAttribute(slice_rows_tiled_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(slice_rows_tiled_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
        const long long ntiles = (Rpad / SL_ROWS) * (K / OZ_KC);
        const unsigned grid = (unsigned)(ntiles < 3ll * sms ? ntiles : 3ll * sms);
        if (one) return launch_pdl(slice_rows_tiled_kernel<S, true>, dim3(grid), dim3(SL_THREADS), smem, st, A0, ld0, K0, A1, ld1, R, K / OZ_KC, Xs, rowscale, chunk_stride);
        return launch_pdl(slice_rows_tiled_kernel<S, false>, dim3(grid), dim3(SL_THREADS), smem, st, A0, ld0, K0, A1, ld1, R, K / OZ_KC, Xs, rowscale, chunk_stride);
    }
    const long long threads = Rpad * (K / 16);
    slice_rows_kernel<S><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(A0, ld0, K0, A1, ld1, K1, R, Xs, rowscale, chunk_stride, slice_onefma());
    return cudaGetLastError();
}

cudaError_t launch_slice_rows(const double* A0, int ld0, int K0, const double* A1, int ld1, int K1, int R, int S,
                              int8_t* Xs, double* rowscale, cudaStream_t st) {
    if (R <= 0) return cudaSuccess;
    const int K = K0 + K1;
    if ((K0 % OZ_KC) != 0 || (K1 % OZ_KC) != 0 || (K != 128 && K != 256 && K != 512)) return cudaErrorInvalidValue;
    cudaError_t e;
    const size_t chunk_stride = ozaki_slices_bytes(R, OZ_KC, S);     // chunk c of the input goes to Xs + c * chunk_stride
    switch (S) {
        case 4: e = slice_rows_t<4>(A0, ld0, K0, A1, ld1, K1, R, Xs, rowscale, chunk_stride, st); break;
        case 5: e = slice_rows_t<5>(A0, ld0, K0, A1, ld1, K1, R, Xs, rowscale, chunk_stride, st); break;
        case 6: e = slice_rows_t<6>(A0, ld0, K0, A1, ld1, K1, R, Xs, rowscale, chunk_stride, st); break;
        case 7: e = slice_rows_t<7>(A0, ld0, K0, A1, ld1, K1, R, Xs, rowscale, chunk_stride, st); break;
        default: return cudaErrorInvalidValue;
    }
    if (e == cudaSuccess) count_launch();
    return e;
}

template <int S, int EPI, int DBG, bool FULL, int GROUPS>
static cudaError_t ozaki_gemm_tg(const OzParams& p, dim3 grid, cudaStream_t st) {
    const size_t smem = (size_t)S * (OZ_XTILE + OZ_WSTAGES * OZ_WTILE);
    cudaError_t e = cudaFuncSetAttribute(ozaki_gemm_kernel<S, EPI, DBG, FULL, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_pdl(ozaki_gemm_kernel<S, EPI, DBG, FULL, GROUPS>, grid, dim3(OZ_THREADS), smem, st, p);
}
template <int S, int EPI, int DBG, bool FULL>
static cudaError_t ozaki_gemm_tf(const OzParams& p, dim3 grid, cudaStream_t st) {
    // MDGAT_OZ_GROUPS=1|2 (read once): epilogue organisation, see the kernel
    static const int groups = [] { const char* e = getenv("MDGAT_OZ_GROUPS"); return e && e[0] == '1' ? 1 : 2; }();
    // two groups need the k chunks of a column tile in the same group (the partial sum is read back by the thread
    // that wrote it): one chunk, or an even number of column tiles per CTA
    // (persistent mode: every segment of a CTA has an even number of column tiles iff items_per_cta and the column
    // tile count are even)
    const bool ok2 = p.K == OZ_KC || ((p.col_tiles_per_cta % 2) == 0 && (p.items_per_cta % 2) == 0);
    return groups == 2 && ok2 ? ozaki_gemm_tg<S, EPI, DBG, FULL, 2>(p, grid, st) : ozaki_gemm_tg<S, EPI, DBG, FULL, 1>(p, grid, st);
}
template <int S, int EPI, int DBG>
static cudaError_t ozaki_gemm_te(const OzParams& p, dim3 grid, cudaStream_t st) {
    // row count a multiple of the tile height (cfg2: 32768 rows): the epilogue drops every row check
    return (p.R % OZ_BM) == 0 ? ozaki_gemm_tf<S, EPI, DBG, true>(p, grid, st) : ozaki_gemm_tf<S, EPI, DBG, false>(p, grid, st);
}
template <int S>
static cudaError_t ozaki_gemm_t(const OzParams& p, dim3 grid, cudaStream_t st) {
    // timeline probes / debug switches live in their own instantiations: 1 = loader + MMA thread only (the epilogue
    // is the production code), 2 = epilogue probes and switches too (costs registers there)
    if ((p.dbg & 255) != 0) return p.epi == EPI_QKV ? ozaki_gemm_te<S, EPI_QKV, 2>(p, grid, st) : ozaki_gemm_te<S, EPI_PLAIN, 2>(p, grid, st);
    if (p.trace != nullptr) return p.epi == EPI_QKV ? ozaki_gemm_te<S, EPI_QKV, 1>(p, grid, st) : ozaki_gemm_te<S, EPI_PLAIN, 1>(p, grid, st);
    return p.epi == EPI_QKV ? ozaki_gemm_te<S, EPI_QKV, 0>(p, grid, st) : ozaki_gemm_te<S, EPI_PLAIN, 0>(p, grid, st);
}

cudaError_t launch_ozaki_gemm(const OzGemmArgs& a, int S, cudaStream_t st) {
    if (a.R <= 0) return cudaSuccess;
    if ((a.K != OZ_KC && a.K != 2 * OZ_KC) || (a.Nout % OZ_BN) != 0 || a.Nout > OZ_MAXN) return cudaErrorInvalidValue;
    if (a.relu && a.Res == a.Y && a.Res != nullptr && a.K > OZ_KC) return cudaErrorInvalidValue;   // see res_first
    if (a.epi == EPI_QKV && a.K != OZ_KC) return cudaErrorInvalidValue;
    // the epilogue addresses Y / Res / the head-major q,k,v rows with 32-bit element offsets
    const long long rpad = (long long)((a.R + OZ_BM - 1) / OZ_BM) * OZ_BM;
    if (rpad * (a.ldy > a.ldres ? a.ldy : a.ldres) >= (1ll << 31) || rpad * HEADS * LDH_QK >= (1ll << 31)) return cudaErrorInvalidValue;
    OzParams p;
    p.Xs[0] = a.Xs[0]; p.Xs[1] = a.Xs[1]; p.rowscale[0] = a.rowscale[0]; p.rowscale[1] = a.rowscale[1]; p.Ws = a.Ws; p.colscale = a.colscale; p.bias = a.bias;
    p.Res = a.Res; p.ldres = a.ldres; p.Y = a.Y; p.ldy = a.ldy; p.R = a.R; p.Nout = a.Nout; p.K = a.K;
    p.relu = a.relu; p.epi = a.epi; p.Qh = a.Qh; p.Kh = a.Kh; p.Vh = a.Vh; p.rows0 = a.rows0; p.n0 = a.n0; p.n1 = a.n1;
    p.trace = g_trace_dev; p.dbg = g_debug_flags;
    p.slice_out = a.slice_out; p.slice_scale = a.slice_scale; p.slice_chunk_stride = ozaki_slices_bytes(a.R, OZ_KC, S);
    p.slice_onefma = slice_onefma() ? 1 : 0;
    const int row_tiles = (a.R + OZ_BM - 1) / OZ_BM, col_tiles = a.Nout / OZ_BN;
    // the X slice planes are loaded once per (CTA, k chunk): keep all column tiles in one CTA unless that leaves
    // SMs idle
    int groups = 1;
    while (row_tiles * groups < 148 && groups < col_tiles) ++groups;
    p.col_tiles_per_cta = (col_tiles + groups - 1) / groups;
    dim3 grid(row_tiles, (col_tiles + p.col_tiles_per_cta - 1) / p.col_tiles_per_cta);
    if (a.slice_out != nullptr && (grid.y != 1 || a.epi != EPI_PLAIN || (a.Nout % OZ_KC) != 0)) return cudaErrorInvalidValue;   // see ozaki_gemm_can_slice
    // Persistent mode: more row tiles than SMs but not a multiple of them -> one CTA per SM walks an equal share of the
    // (row tile, column tile) items (MDGAT_OZ_PERSISTENT=0 keeps one CTA per row tile).
    p.items_per_cta = 0;
    static const bool persistent_env = [] { const char* e = getenv("MDGAT_OZ_PERSISTENT"); return !(e && e[0] == '0'); }();
    static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
    if (persistent_env && grid.y == 1 && row_tiles > sms && (row_tiles % sms) != 0 && a.slice_out == nullptr) {
        const int total = row_tiles * col_tiles;
        p.items_per_cta = (total + sms - 1) / sms;
        // MDGAT_OZ_EVEN_ITEMS=1: with two k chunks, round the share up to an even count so that the two-group epilogue applies
        // (256 -> 128 at cfg2: 8 instead of 7 items, 128 instead of 147 CTAs)
        static const bool even_env = [] { const char* e = getenv("MDGAT_OZ_EVEN_ITEMS"); return e && e[0] == '1'; }();
        if (even_env && a.K > OZ_KC && (p.items_per_cta & 1) && (col_tiles % 2) == 0) ++p.items_per_cta;
        grid = dim3((total + p.items_per_cta - 1) / p.items_per_cta, 1);
    }
    cudaError_t e;
    switch (S) {
        case 4: e = ozaki_gemm_t<4>(p, grid, st); break;
        case 5: e = ozaki_gemm_t<5>(p, grid, st); break;
        case 6: e = ozaki_gemm_t<6>(p, grid, st); break;
        case 7: e = ozaki_gemm_t<7>(p, grid, st); break;
        default: return cudaErrorInvalidValue;
    }
    if (e == cudaSuccess) count_launch();
    return e;
}

}  // namespace mdgat
