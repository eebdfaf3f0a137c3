// float64-faithful multi-head attention on the 5th-generation tensor cores: Q K^T and P V of
// attention() / dynamic_attention() (/root/reference/models/mdgat.py:190-210) as exact int8 slice
// products (tcgen05.mma kind::i8, int32 accumulators in TMEM, operands staged by TMA bulk copies).
//
// tcgen05 has no f64 MMA kind, and the DMMA flash kernel (attention_f64.cu) spends 64 FP64-pipe FLOPs per
// logit on Q K^T and 64 on P V. Here both contractions leave the FP64 pipe:
//
//  * slice_qk_kernel / slice_v_kernel write every head vector as S balanced base-256 digits,
//        x = 2^(e-8S+2) * sum_{s=0..S-1} D_s 256^(S-1-s),   D_s in [-128, 127],
//    e = exponent of the row maximum for q and k rows, of the column (channel) maximum over all source
//    keypoints for v. S digits = 8S-1 bits (S = 5: 39 bits, the default chosen on the parity sweep; S = 7: 55
//    bits, "exact" mode). Planes are laid out in the UMMA canonical no-swizzle K-major order, one contiguous
//    block per tile, so one cp.async.bulk brings a tile in.
//  * Q K^T: the S(S+1)/2 digit products D_s G_t^T with s + t <= S-1 are accumulated exactly in int32; products
//    with the same s + t share a TMEM column group ("diagonal"). Stacked-N issue: plane s of Q against planes
//    0..S-1-s of K in ONE MMA of N = (S-s)*32 written 32*s columns into the accumulator set. With S <= 5 there
//    are TWO accumulator sets, so the tensor core multiplies tile j+1 while the epilogue recombines tile j.
//    The epilogue recombines the diagonals in float64 (Horner, neighbouring diagonals merged exactly in int32).
//  * softmax needs exp(z - max) <= 1 before P can be cut into digits, and an integer accumulator cannot be
//    rescaled when a running maximum moves. So the row maximum comes first: PASS 1 multiplies only the top
//    two diagonals (3 digit products) and takes the row maximum in fp32; a rigorous bound on what the
//    dropped digits can add turns it into c_i >= max_j z_ij with c_i - max <~ 1-2 (1-3 bits of P).
//  * PASS 2: p = exp(z - c_i) in float64 (256-entry table + cubic / quartic), p^ = rint(p 2^(8 SP - 1)) read
//    straight out of the mantissa as SP unsigned bytes = the SP digit planes of P (4x4 byte transposes), written
//    to shared memory in A-operand order. P V: unsigned P digits x signed V digits, the products with
//    a + t <= S-1, accumulated over ALL source keypoints in TMEM (no online rescaling), one Horner pass per query
//    row at the end, divided by the exact integer row sum of the p^ (per-plane byte sums by DP4A).
//
// LOGITS mode (dynamic layers, attention = 'tcgen05_i8_all') stops after the Horner pass and stores the scaled
// logits for the exact top-k selection kernel.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int AI_BM = 128;                 // query rows per CTA = TMEM lanes
constexpr int AI_BN = 32;                  // source keypoints per tile
constexpr int AI_QPLANE = AI_BM * 32;      // bytes of one Q digit plane of a query tile
constexpr int AI_KPLANE = AI_BN * 32;      // bytes of one K (or V^T) digit plane of a source tile
constexpr int AI_PPLANE = AI_BM * AI_BN;   // bytes of one P digit plane
constexpr int AI_STAGES = 4;               // K/V tile ring (the logits run two tiles ahead of P V)
// epilogue organisation: CW = columns of a 32-column tile per warp (16: 8 epilogue warps, 8: 16 epilogue warps = 4 per SM
// sub-partition); the MMA and loader warps follow the epilogue warps
constexpr int ai_epi_threads(int cw) { return 128 * (32 / cw); }
constexpr int ai_threads(int cw) { return ai_epi_threads(cw) + 64; }
constexpr int AI_CVT_DEFAULT = 4;          // int32 -> float64 conversion of the epilogue (int_to_f64)
constexpr int AI_EXP_LIMIT = 60;           // |exponent| clamp of the digit scales (values beyond 2^60 are out of range)

size_t attn_i8_q_bytes(int B, int n, int S) { return (size_t)B * HEADS * ((n + AI_BM - 1) / AI_BM) * S * AI_QPLANE; }
size_t attn_i8_kv_bytes(int B, int n, int S) { return (size_t)B * HEADS * ((n + AI_BN - 1) / AI_BN) * S * AI_KPLANE; }
static int pad_to(int n, int a) { return (n + a - 1) / a * a; }
static int tiles_pad4(int n) { return (((n + AI_BN - 1) / AI_BN) + 3) & ~3; }     // ktilemax row stride: 16-byte bulk copies

size_t attn_i8_side_bytes(int B, int n, int S) {
    // Q planes | K planes | V planes | qscale[B*4*npad128] | kscale_d[B*4*npad32] | vscale[B*4*32] | kscale_f | ktilemax
    const size_t rq = (size_t)B * HEADS * pad_to(n, AI_BM), rk = (size_t)B * HEADS * pad_to(n, AI_BN);
    size_t b = attn_i8_q_bytes(B, n, S) + 2 * attn_i8_kv_bytes(B, n, S);
    b += rq * 8 + rk * 8 + (size_t)B * HEADS * 32 * 8 + rk * 4 + (size_t)B * HEADS * tiles_pad4(n) * 4;
    return (b + 255) / 256 * 256;
}

AttnI8Side attn_i8_carve(void* base, int B, int n, int S) {
    AttnI8Side s;
    const size_t rq = (size_t)B * HEADS * pad_to(n, AI_BM), rk = (size_t)B * HEADS * pad_to(n, AI_BN);
    unsigned char* p = reinterpret_cast<unsigned char*>(base);
    s.Qs = reinterpret_cast<int8_t*>(p); p += attn_i8_q_bytes(B, n, S);
    s.Ks = reinterpret_cast<int8_t*>(p); p += attn_i8_kv_bytes(B, n, S);
    s.Vs = reinterpret_cast<int8_t*>(p); p += attn_i8_kv_bytes(B, n, S);
    s.qscale = reinterpret_cast<double*>(p); p += rq * 8;
    s.kexp = reinterpret_cast<int*>(p); p += rk * 8;                 // rk * 4 used
    s.vscale = reinterpret_cast<double*>(p); p += (size_t)B * HEADS * 32 * 8;
    s.kscale_f = reinterpret_cast<float*>(p); p += rk * 4;
    s.ktilemax = reinterpret_cast<float*>(p);
    s.n = n;
    s.S = S;
    return s;
}

DEVINL double pow2i(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// canonical K-major core-matrix offset inside a plane whose rows hold 32 bytes of K: 8 rows x 16 bytes contiguous,
// the two 16-byte K halves 128 B apart (LBO), 8-row groups 256 B apart (SBO)
DEVINL int canon32(int r, int khalf) { return (r >> 3) * 256 + khalf * 128 + (r & 7) * 16; }

// S balanced base-256 digits of 16 values -> w[plane][4] (16 bytes per plane, plane 0 = most significant).
// I = rint(x sc), |I| <= 65.5 * 256^(S-1). Adding the bias B = 0x80 in every byte position makes I + B non-negative, and the
// balanced digits are then simply its bytes with the top bit flipped (sum (u_s - 128) 256^k = I, no carries). For S <= 6 the
// biased integer fits the mantissa: ONE fma(x, sc, 1.5 * 2^52 + B) leaves it in the low mantissa bits -- one FP64 instruction,
// two LOP3 and a 4x4 byte transpose per value instead of a 64-bit conversion and a shift / subtract chain per digit.
template <int S>
DEVINL void digits16(const double* x, double sc, uint32_t (&w)[S][4]) {
    if constexpr (S <= 6) {
        constexpr unsigned long long BIAS = S == 4 ? 0x80808080ull : (S == 5 ? 0x8080808080ull : 0x808080808080ull);
        const double magic = 6755399441055744.0 + (double)BIAS;        // exact: BIAS < 2^48
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double u = fma(x[4 * g + i], sc, magic);
                lo[i] = (uint32_t)__double2loint(u) ^ 0x80808080u;          // bytes 0..3: digits S-1 .. S-4
                hi[i] = (uint32_t)__double2hiint(u) ^ 0x8080u;              // bytes 0..1: digits S-5, S-6 (bits 32..47 of the integer)
            }
            const uint32_t a0 = __byte_perm(lo[0], lo[1], 0x5140), a1 = __byte_perm(lo[0], lo[1], 0x7362);
            const uint32_t b0 = __byte_perm(lo[2], lo[3], 0x5140), b1 = __byte_perm(lo[2], lo[3], 0x7362);
            w[S - 1][g] = __byte_perm(a0, b0, 0x5410);
            w[S - 2][g] = __byte_perm(a0, b0, 0x7632);
            w[S - 3][g] = __byte_perm(a1, b1, 0x5410);
            w[S - 4][g] = __byte_perm(a1, b1, 0x7632);
            if (S > 4) {
                const uint32_t c0 = __byte_perm(hi[0], hi[1], 0x5140), d0 = __byte_perm(hi[2], hi[3], 0x5140);
                w[S > 4 ? S - 5 : 0][g] = __byte_perm(c0, d0, 0x5410);
                if (S > 5) w[S > 5 ? S - 6 : 0][g] = __byte_perm(c0, d0, 0x7632);
            }
        }
        return;
    }
#pragma unroll
    for (int s = 0; s < S; ++s) { w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u; }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        long long I = __double2ll_rn(x[i] * sc);           // |I| <= 2^(8S-2)
#pragma unroll
        for (int s = S - 1; s >= 1; --s) {
            const int d = (int)(((unsigned)I & 0xffu) ^ 0x80u) - 0x80;     // low byte as a signed digit
            I = (I - d) >> 8;
            w[s][i >> 2] |= (uint32_t)(d & 0xff) << (8 * (i & 3));
        }
        w[0][i >> 2] |= (uint32_t)((int)I & 0xff) << (8 * (i & 3));        // |top digit| <= 65
    }
}

// ---------------------------------------------------------------------------------------------------
// Digits of the q and k rows. One thread per (b, h, padded row); blocks [0, nqb) cut Q (query tiles of 128
// rows), the rest cut K (4 source tiles of 32 rows per block).
// qscale = 2^(e-12) / sqrt(32)   (row factor of the logits, 1/sqrt(d) of mdgat.py:192 included)
// kexp = f << 20 (added to the exponent field of the row factor: 2^f), kscale_f = 2^f as float, ktilemax = largest kscale_f of a 32-row tile
// ---------------------------------------------------------------------------------------------------
// blk128 / t128: index of the 128-row block and the thread inside it (the stand-alone kernel maps them to blockIdx /
// threadIdx, the fused per-layer kernel packs four of them into a 512-thread CTA)
// STAGED (the forward path): every thread first brings ITS row (256 bytes) into shared memory with one TMA bulk copy, at a
// 272-byte pitch (16 bytes past a multiple of 128: the LDS.128 of 32 consecutive rows are conflict-free). Read straight
// from global memory, a warp's 16-byte loads sit 288 bytes apart -- 32 cache lines per instruction, 48 of them per row --
// and the kernel was bound by the L1 wavefront pipe (53 % busy, lg_throttle its largest stall). stage: this warp's 32 x 272
// bytes; bar: this warp's mbarrier (initialised by the caller, one arrival).
constexpr int SQ_PITCH = 272;
template <int S, bool STAGED>
DEVINL void slice_qk_body(const double* __restrict__ Qh, const double* __restrict__ Kh, const AttnI8Side& o, int nqb, int blk128, int t128,
                          unsigned char* stage = nullptr, uint64_t* bar = nullptr) {
    const bool isq = blk128 < nqb;
    const int n = o.n;
    const int npad = isq ? (n + AI_BM - 1) / AI_BM * AI_BM : (n + AI_BN - 1) / AI_BN * AI_BN;
    const int bpb = (npad + 127) / 128;                              // blocks per (b, h)
    const int blk = isq ? blk128 : blk128 - nqb;
    const int bh = blk / bpb;
    const int i = (blk - bh * bpb) * 128 + t128;                     // padded row inside (b, h)
    if (i >= npad) return;                                           // K only; whole warps (npad multiple of 32)
    const double* src = (isq ? Qh : Kh) + ((long long)bh * n + i) * LDH_QK;
    if (STAGED) {
        const unsigned have = __ballot_sync(0xffffffffu, i < n);
        if (have) {
            if ((t128 & 31) == 0) mbar_expect_tx(bar, (unsigned)__popc(have) * 256u);
            __syncwarp();
            unsigned char* mine = stage + (t128 & 31) * SQ_PITCH;
            if (i < n) bulk_g2s(mine, src, 256, bar);
            mbar_wait(bar, 0);
            src = reinterpret_cast<const double*>(mine);
        }
    }
    // 16-byte read of the row at element c: LDS from the staged copy, else a global load
    const uint32_t src_s32 = STAGED ? (uint32_t)__cvta_generic_to_shared(src) : 0u;
    auto ld2 = [&](int c) -> double2 {
        if (STAGED) {
            double2 v;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(src_s32 + (uint32_t)c * 8u));
            return v;
        }
        return *reinterpret_cast<const double2*>(src + c);
    };
    // the row is read twice (maximum, then 16 values at a time for the digits; the second read hits L1): 16 instead of
    // 32 live doubles keep the fused slicer at two 512-thread CTAs per SM
    double mx = 0.0;
    if (i < n) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) { const double2 v = ld2(c); mx = fmax(mx, fmax(fabs(v.x), fabs(v.y))); }
    }
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);                                     // |x| < 2^e
    e = max(-AI_EXP_LIMIT, min(AI_EXP_LIMIT, e));
    const double sc = pow2i(8 * S - 2 - e);
    int8_t* dst;
    int plane;
    if (isq) {
        dst = o.Qs + ((size_t)bh * (npad / AI_BM) + (i >> 7)) * (S * AI_QPLANE) + canon32(i & 127, 0);
        plane = AI_QPLANE;
        o.qscale[(size_t)bh * npad + i] = i < n ? pow2i(e - 12) * 0.17677669529663688110 : 0.0;
    } else {
        dst = o.Ks + ((size_t)bh * (npad / AI_BN) + (i >> 5)) * (S * AI_KPLANE) + canon32(i & 31, 0);
        plane = AI_KPLANE;
        const float kf = i < n ? __int_as_float((127 + e) << 23) : 0.f;
        o.kexp[(size_t)bh * npad + i] = i < n ? e * (1 << 20) : 0;
        o.kscale_f[(size_t)bh * npad + i] = kf;
        float tm = kf;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, off));
        if ((t128 & 31) == 0) o.ktilemax[(size_t)bh * (((npad / AI_BN) + 3) & ~3) + (i >> 5)] = tm;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        double x[16];
        if (i < n) {
#pragma unroll
            for (int c = 0; c < 16; c += 2) { const double2 v = ld2(half * 16 + c); x[c] = v.x; x[c + 1] = v.y; }
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = 0.0;
        }
        uint32_t w[S][4];
        digits16<S>(x, sc, w);
#pragma unroll
        for (int s = 0; s < S; ++s)
            *reinterpret_cast<uint4*>(dst + (size_t)s * plane + half * 128) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

template <int S>
__global__ void __launch_bounds__(128)
slice_qk_kernel(const double* __restrict__ Qh, const double* __restrict__ Kh, AttnI8Side o, int B, int nqb) {
    slice_qk_body<S, false>(Qh, Kh, o, nqb, (int)blockIdx.x, (int)threadIdx.x);
}

// ---------------------------------------------------------------------------------------------------
// Digits of V, transposed: the P V product contracts over source keypoints, so the B operand is V^T
// (32 channel rows x 32 keypoints of K per tile) and all keypoints of a (b, h) share one exponent per
// channel. One CTA per (b, h); lane = channel. vscale[c] = 2^(e_c - 13): with p^ = p 2^(8 SP - 1) and the digit
// weights, message = vscale * Horner(P V diagonals) * 2^(8 SP - 1) / sum_j p^_j.
// ---------------------------------------------------------------------------------------------------
constexpr int SV_WARPS = 16;     // one CTA per (b, h) (128 CTAs at cfg2): 16 warps keep enough loads in flight per SM
template <int S>
DEVINL void slice_v_body(const double* __restrict__ Vh, const AttnI8Side& o, int bh) {
    __shared__ double s_max[SV_WARPS][32];
    const int n = o.n, npad = (n + AI_BN - 1) / AI_BN * AI_BN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* V = Vh + (long long)bh * n * LDH_V;
    double mx = 0.0;
    for (int j = warp; j < n; j += SV_WARPS) mx = fmax(mx, fabs(V[(long long)j * LDH_V + lane]));
    s_max[warp][lane] = mx;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < SV_WARPS; ++w) mx = fmax(mx, s_max[w][lane]);
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);
    e = max(-AI_EXP_LIMIT, min(AI_EXP_LIMIT, e));
    if (warp == 0) o.vscale[(size_t)bh * 32 + lane] = pow2i(e - 13);
    const double sc = pow2i(8 * S - 2 - e);
    // unit = 16 consecutive keypoints: 16 bytes per plane and channel
    for (int unit = warp; unit < npad / 16; unit += SV_WARPS) {
        const int j0 = unit * 16;
        double x[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) x[jj] = (j0 + jj) < n ? V[(long long)(j0 + jj) * LDH_V + lane] : 0.0;
        uint32_t w[S][4];
        digits16<S>(x, sc, w);
        int8_t* dst = o.Vs + ((size_t)bh * (npad / AI_BN) + (j0 >> 5)) * (S * AI_KPLANE) + canon32(lane, (j0 >> 4) & 1);
#pragma unroll
        for (int s = 0; s < S; ++s)
            *reinterpret_cast<uint4*>(dst + (size_t)s * AI_KPLANE) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

template <int S>
__global__ void __launch_bounds__(32 * SV_WARPS)
slice_v_kernel(const double* __restrict__ Vh, AttnI8Side o) { slice_v_body<S>(Vh, o, (int)blockIdx.x); }

// The forward path: both sides of a layer in TWO launches -- the q / k rows (128-thread CTAs, one 128-row block each, rows
// staged through shared memory) and the v channels (512-thread CTAs, one per (b, h)).
struct SliceSide { const double *Qh, *Kh, *Vh; AttnI8Side o; int nqb, nkb, qk_ctas, v_ctas; };
template <int S>
__global__ void __launch_bounds__(128, 6)
slice_qk_sides_kernel(const __grid_constant__ SliceSide s0, const __grid_constant__ SliceSide s1) {
    extern __shared__ __align__(128) unsigned char sq_smem[];           // [4 warps][32 rows][272 B]
    __shared__ __align__(8) uint64_t bars[4];
    const int warp = (int)threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) mbar_init(&bars[warp], 1);
    mbar_fence_init();
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    int blk = blockIdx.x;
    const bool second = blk >= s0.nqb + s0.nkb;
    const SliceSide& sd = second ? s1 : s0;
    if (second) blk -= s0.nqb + s0.nkb;
    slice_qk_body<S, true>(sd.Qh, sd.Kh, sd.o, sd.nqb, blk, (int)threadIdx.x, sq_smem + warp * 32 * SQ_PITCH, &bars[warp]);
}
template <int S>
__global__ void __launch_bounds__(32 * SV_WARPS, 2)
slice_v_sides_kernel(const __grid_constant__ SliceSide s0, const __grid_constant__ SliceSide s1) {
    pdl_wait();
    pdl_trigger();
    int blk = blockIdx.x;
    const bool second = blk >= s0.v_ctas;
    const SliceSide& sd = second ? s1 : s0;
    if (second) blk -= s0.v_ctas;
    slice_v_body<S>(sd.Vh, sd.o, blk);
}

// ---------------------------------------------------------------------------------------------------
// tcgen05 helpers (raw PTX)
// ---------------------------------------------------------------------------------------------------
DEVINL uint64_t ai_desc(const void* smem) {      // K-major, no swizzle, LBO 128 B, SBO 256 B
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
DEVINL void ai_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
DEVINL void ai_commit(uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(b) : "memory");
}
DEVINL void ai_ld16(uint32_t taddr, int (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),
                   "=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
                 : "r"(taddr));
}
DEVINL void ai_ld8(uint32_t taddr, int (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(taddr));
}
DEVINL void ai_ldw(uint32_t taddr, int (&r)[16]) { ai_ld16(taddr, r); }
DEVINL void ai_ldw(uint32_t taddr, int (&r)[8]) { ai_ld8(taddr, r); }
DEVINL void ai_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
DEVINL void ai_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DEVINL void ai_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int THREADS> DEVINL void epi_bar_sync() { asm volatile("bar.sync 1, %0;" :: "n"(THREADS) : "memory"); }   // the epilogue warps only
// int32 -> float64, exact. CVT selects the instruction mix (the epilogue is issue-bound, the three pipes are not):
//   0  sign bit flipped into the low mantissa word of 2^52 (LOP3 + a MOV for the high word), minus 2^52 + 2^31 (DADD)
//   1  I2F.F64.S32: one instruction on the conversion pipe (16 lanes / clk / SM), nothing on the FP64 pipe
//   2  the bit pattern of 2^52 + 2^31 + v built by 64-bit integer multiply-adds (IMAD.WIDE), minus 2^52 + 2^31 (DADD)
constexpr long long AI_MAGIC_BITS = 0x4330000080000000LL;      // 2^52 + 2^31 as a double
template <int CVT>
DEVINL double int_to_f64(int v) {
    if (CVT == 1) return __int2double_rn(v);
    if (CVT == 2) {
        long long t;
        asm("mad.wide.s32 %0, %1, 1, %2;" : "=l"(t) : "r"(v), "l"(AI_MAGIC_BITS));
        return __longlong_as_double(t) - 4503601774854144.0;
    }
    return __hiloint2double(0x43300000, v ^ (int)0x80000000) - 4503601774854144.0;
}
// a * 256 + b for |a|, |b| < 2^22 (two neighbouring diagonals), as a float64
template <int CVT>
DEVINL double pair_to_f64(int a, int b) {
    if (CVT == 2) {
        long long t, u;
        asm("mad.wide.s32 %0, %1, 1, %2;" : "=l"(t) : "r"(b), "l"(AI_MAGIC_BITS));
        asm("mad.wide.s32 %0, %1, 256, %2;" : "=l"(u) : "r"(a), "l"(t));
        return __longlong_as_double(u) - 4503601774854144.0;
    }
    return int_to_f64<CVT>(a * 256 + b);
}
// instruction descriptor: D = s32, B = signed 8-bit, both operands K-major, M = 128; A signed or unsigned
DEVINL constexpr uint32_t ai_idesc(int n, bool a_signed) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(AI_BM >> 4) << 24);
}

// sum_dd acc[dd][j] 256^(-dd), times 256 when S is even (the caller folds 2^-8 into the row factor): the diagonals are
// merged in pairs exactly in int32 (|acc_dd| <= S * 32 * 2^14 < 2^22) from the least significant end, then Horner in float64
// CVT == 3: the least significant pair through I2F (conversion pipe), the others through the 2^52 constant (ALU + FP64 pipe)
// CVT == 4 (S <= 6): all diagonals merged into ONE 64-bit integer (|sum| < 2^63: S = 5 needs 54 bits) by integer
//           multiply-adds, then a single I2F.F64.S64 -- no FP64-pipe instruction at all; the result is 2^(16 (S-1)/2) times
//           the Horner value (ai_recombine_scale), which the caller folds into the row factor
template <int S, int CVT> constexpr double ai_recombine_scale() {
    return (CVT == 4 && S <= 6) ? 1.0 / (double)(1ull << (16 * ((S - 1) / 2))) : 1.0;
}
template <int S, int CW, int CVTX>
DEVINL double ai_recombine(const int (&acc)[S][CW], int j) {
    if constexpr (CVTX == 4 && S <= 6) {
        long long t = (long long)(acc[S - 2][j] * 256 + acc[S - 1][j]);
        int sh = 16;
#pragma unroll
        for (int d = S - 4; d >= (S & 1); d -= 2) {
            const int g = acc[d][j] * 256 + acc[d + 1][j];
            if (sh == 16) asm("mad.wide.s32 %0, %1, 65536, %0;" : "+l"(t) : "r"(g));
            else t += (long long)g << sh;
            sh += 16;
        }
        if (S & 1) t += (long long)acc[0][j] << sh;
        return __ll2double_rn(t);
    }
    constexpr int CVT = (CVTX == 3 || CVTX == 4) ? 0 : CVTX;
    double h = pair_to_f64<CVTX == 3 ? 1 : CVT>(acc[S - 2][j], acc[S - 1][j]);
#pragma unroll
    for (int d = S - 4; d >= (S & 1); d -= 2) h = fma(h, 1.52587890625e-05, pair_to_f64<CVT>(acc[d][j], acc[d + 1][j]));
    if (S & 1) h = fma(h, 1.52587890625e-05, int_to_f64<CVT>(acc[0][j]));
    return h;
}

// rint(exp(x) 2^(8 SP - 1)) for x <= 0 as a fixed-point integer read out of the mantissa: lo = bits 0..31, hi = bits 32..47.
// 256-entry table 2^(j/256), |r| <= ln2/512: the cubic is good to 1.4e-13 relative (SP <= 4: 32-bit probabilities), the
// quartic to 4e-17 (SP > 4); one correctly rounded ln2/256 suffices under the FMA. The power of two, the 2^(8 SP - 1)
// scale and the underflow clamp are one integer add into the exponent field; results below 1/2 round to 0.
// Table of 2^(j/E): E = 16 entries for SP <= 4 -- 128 bytes, one entry per bank pair, so the 32 random lookups of a warp
// never conflict (a 256-entry table costs ~9 shared-memory wavefronts per warp instead of 2); |r| <= ln2/32 and the quartic
// is good to 4e-11 relative, a fifth of the last probability bit. SP > 4 (float64-faithful mode): 256 entries, |r| <= ln2/512,
// quartic good to 4e-17. tbl_s32: shared-window address of the table, aligned to its size (entry address = base | index * 8).
template <int SP>
DEVINL void ai_exp_fixed(double x, uint32_t tbl_s32, uint32_t& lo, uint32_t& hi) {
    constexpr int E = SP > 4 ? 256 : 16, EL = SP > 4 ? 8 : 4;
    const double MAGIC = 6755399441055744.0;                // 1.5 * 2^52: rint() in the low mantissa bits
    // n + E (8 SP - 1) in the low mantissa word: the 2^(8 SP - 1) scale rides along in the exponent part n >> EL
    const double MAGIC_N = MAGIC + (double)E * (8 * SP - 1);
    const double tn = fma(x, EXP_INV_LN2_256 * (E / 256.0), MAGIC_N);
    const int n = __double2loint(tn);
    const double nd = tn - MAGIC_N;
    const double r = fma(nd, -EXP_LN2_256 * (256.0 / E), x);
    const double r2 = r * r;
    double q = fma(r, 1.0 / 24.0, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    const double pl = fma(r2, q, r);                         // e^r - 1
    double t;
    asm("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(tbl_s32 | (((uint32_t)n << 3) & (uint32_t)(8 * E - 8))));
    const double y = fma(t, pl, t);                          // in (0.97, 2.05)
    const int m = max(n >> EL, -9);                          // exp(x) 2^(8 SP - 1) < 2^-9 rounds to 0 like anything smaller
    const double ys = __hiloint2double(__double2hiint(y) + (m << 20), __double2loint(y));
    const double pm = ys + MAGIC;
    lo = (uint32_t)__double2loint(pm);
    hi = (uint32_t)__double2hiint(pm) & 0xffffu;
}

// The same fixed-point exponential with the logit's own multiply-add folded in (FULL mode, SP <= 4; MDGAT_ATTN_CVT=5|6).
// ai_exp_fixed() is fed x = z rk - c_i (one DFMA) and spends a second DFMA on x E/ln2 + MAGIC. Here the row factor carries
// E/ln2 (rk16 = rk E/ln2: the key exponent is still one integer add) and the row's bound is rounded UP to a whole number K of
// ln2/E steps (c_i grows by < ln2/E: 0.06 bit of P at E = 16), so that
//     tn = z rk16 + (MAGIC_N - K)          low mantissa word: n = rint(z rk16 - K) + E (8 SP - 1), exactly as before
//     f  = z rk16 - (tn - (MAGIC_N - K))   = (z rk16 - K) - rint(z rk16 - K), |f| <= 1/2, in units of ln2/E
// and e^(f ln2/E) - 1 is a polynomial in f with the powers of ln2/E folded into its coefficients: 9 FP64 instructions per
// logit instead of 10 (E = 16, quartic), 8 with the 64-entry table and the cubic (|f ln2/64| <= 0.0054: 3.6e-11 relative, the
// same fifth of the last probability bit; the larger table costs shared-memory wavefronts, the pipe the epilogue has to spare).
template <int SP, int E>
DEVINL void ai_exp_fixed_z(double z, double rk16, double CA, uint32_t tbl_s32, uint32_t& lo) {
    static_assert(SP <= 4 && (E == 16 || E == 64), "32-bit probabilities only");
    constexpr int EL = E == 64 ? 6 : 4;
    constexpr double L = 0.6931471805599453 / E;
    const double MAGIC = 6755399441055744.0;
    const double tn = fma(z, rk16, CA);
    const int n = __double2loint(tn);
    const double nd = tn - CA;                               // rint(z rk16 - K) + K: exact (integers below 2^53)
    const double f = fma(z, rk16, -nd);
    double q = E == 64 ? L * L * L / 6.0 : fma(f, L * L * L * L / 24.0, L * L * L / 6.0);
    q = fma(q, f, L * L / 2.0);
    q = fma(q, f, L);
    const double pl = q * f;                                 // e^(f L) - 1
    double t;
    asm("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(tbl_s32 | (((uint32_t)n << 3) & (uint32_t)(8 * E - 8))));
    const double y = fma(t, pl, t);
    const int m = max(n >> EL, -9);
    const double ys = __hiloint2double(__double2hiint(y) + (m << 20), __double2loint(y));
    lo = (uint32_t)__double2loint(ys + MAGIC);
}

struct AttnI8Params {
    AttnI8Side q[2];             // digit planes of the QUERY side of grid side s
    AttnI8Side kv[2];            // digit planes of its SOURCE side
    double* Out[2];              // messages (rows x ldo) or, LOGITS, dense (B,4,N,M) logits
    AttnI8TopK tk;               // TOPK: per-row threshold / last tied column / maximum (launch_topk_threshold)
    AttnI8MsgPlanes mp;          // Xs != nullptr: messages as GEMM digit planes instead of float64 rows
    int B, ldo;
};

// One CTA = 128 query rows of one (side, b, h), warp specialised:
//   last warp, one lane     loader: TMA bulk copies of the Q planes, the key scales and the K / V^T tile ring
//   last-but-one, one lane  MMA issuer
//   the others              epilogue: thread = query row (TMEM lane = 32 * (warp % 4) + lane); the 32 / CW warps that share
//                           a lane quarter split the 32 columns of a tile (CW = 8: 16 epilogue warps, CW = 16: 8)
// S digit planes of q, k, v; SP byte planes of P. NSBUF accumulator sets for Q K^T (2 when TMEM has room: S <= 5).
// MODE: AI_MODE_FULL (pass 1 + pass 2), AI_MODE_LOGITS (the scaled logits are stored, nothing else), AI_MODE_TOPK (no pass 1:
// the exact row maximum and the top-k threshold come from topk_threshold_kernel; probabilities outside the kept set are 0,
// which turns dynamic_attention() of mdgat.py:196-210 into the same tensor-core P V as the full layers).
template <int S, int SP, int MODE, int CW, int CVTX>
__global__ void __launch_bounds__(ai_threads(CW), 1) attn_i8_kernel(const __grid_constant__ AttnI8Params p) {
    // CVTX 0..4: recombination variant with ai_exp_fixed(); 5 / 6: variant 4 with ai_exp_fixed_z() on 16 / 64 table entries
    constexpr int CVT = CVTX >= 5 ? 4 : CVTX;
    constexpr int XE = (CVTX >= 5 && SP <= 4 && MODE == AI_MODE_FULL) ? (CVTX == 6 ? 64 : 16) : 0;     // 0: ai_exp_fixed()
    constexpr int ETAB = SP > 4 ? 256 : (XE == 64 ? 64 : 16);
    constexpr bool LOGITS = MODE == AI_MODE_LOGITS, TOPK = MODE == AI_MODE_TOPK, PASS1 = MODE == AI_MODE_FULL;
    constexpr int AI_EPI_THREADS = ai_epi_threads(CW), EPI_WARPS = AI_EPI_THREADS / 32, NCG = 32 / CW;
    constexpr int NSBUF = (3 * S * AI_BN <= 512) ? 2 : 1;
    constexpr int TM_S = 0;                               // TMEM columns: NSBUF sets of S logits diagonals
    constexpr int TM_O = NSBUF * S * AI_BN;               //               S P V diagonals
    // pass 1 borrows four 64-column buffers: two at the start of the P V region, two at the start of the LAST logits set
    // (the first Q K^T product into that set waits for the end of pass 1)
    constexpr int TM_P1_LO = TM_O, TM_P1_HI = TM_S + (NSBUF - 1) * S * AI_BN;
    constexpr int STAGE_BYTES = 2 * S * AI_KPLANE;
    constexpr int P1G = STAGE_BYTES / (2 * AI_KPLANE);    // pass-1 tiles per ring stage (S: their two leading K planes fill a stage)
    static_assert(SP <= S && SP >= 3 && SP <= 6 && S * AI_BN >= 128 && TM_O + S * AI_BN <= 512, "plane counts");
    extern __shared__ __align__(128) unsigned char ai_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int side = (int)blockIdx.z >= p.B ? 1 : 0;
    const int b = blockIdx.z - side * p.B, h = blockIdx.y, qt = blockIdx.x;
    const AttnI8Side& Qd = p.q[side];
    const AttnI8Side& Kd = p.kv[side];
    const int N = Qd.n, M = Kd.n;
    if (qt * AI_BM >= N) return;
    const int bh = b * HEADS + h;
    const int T = (M + AI_BN - 1) / AI_BN, Mpad = T * AI_BN;
    const int Npad = (N + AI_BM - 1) / AI_BM * AI_BM;

    // the exponential table comes first, on a 2 KB boundary of the shared window (see ai_exp_fixed)
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(ai_smem);
    unsigned char* smem_al = ai_smem + (((smem0 + 2047u) & ~2047u) - smem0);
    double* etab = reinterpret_cast<double*>(smem_al);                            // [256] 2^(j/256)
    const uint32_t etab_s32 = (uint32_t)__cvta_generic_to_shared(etab);
    int8_t* sQ = reinterpret_cast<int8_t*>(smem_al + 2048);                       // [S][4096]
    int8_t* sKV = sQ + S * AI_QPLANE;                                             // [stage][K S planes | V S planes]
    uint8_t* sP = reinterpret_cast<uint8_t*>(sKV + AI_STAGES * STAGE_BYTES);      // [2][SP][4096]
    int* s_kex = reinterpret_cast<int*>(sP + 2 * SP * AI_PPLANE);                 // [Mpad] key exponents << 20
    float* s_ksf = reinterpret_cast<float*>(s_kex + Mpad);                        // [Mpad]
    float* s_ktm = s_ksf + Mpad;                                                  // [T] (padded to 4)
    double* s_xd = reinterpret_cast<double*>(s_ktm + ((T + 3) & ~3));             // [4][128] row exchange between column groups
    unsigned long long* s_xu = reinterpret_cast<unsigned long long*>(s_xd + 512); // [4][128]
    double* s_lg = reinterpret_cast<double*>(s_xu + 512);                          // LOGITS: [128][33] staging tile (coalesced stores)

    __shared__ __align__(8) uint64_t q_full, kv_full[AI_STAGES], kv_empty[AI_STAGES], s1_full[4], s1_empty[4], p1_done,
        s_full[NSBUF], s_empty[NSBUF], p_full[2], p_empty[2], o_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_vexp[4];                                 // message planes: largest exponent word of the head's value scales

    if (tid == 0) {
        mbar_init(&q_full, 1);
#pragma unroll
        for (int i = 0; i < AI_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
#pragma unroll
        for (int i = 0; i < 4; ++i) { mbar_init(&s1_full[i], 1); mbar_init(&s1_empty[i], AI_EPI_THREADS); }
        mbar_init(&p1_done, AI_EPI_THREADS);
#pragma unroll
        for (int i = 0; i < 2; ++i) { mbar_init(&p_full[i], AI_EPI_THREADS); mbar_init(&p_empty[i], 1); }
#pragma unroll
        for (int i = 0; i < NSBUF; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], AI_EPI_THREADS); }
        mbar_init(&o_full, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < ETAB; i += ai_threads(CW)) etab[i] = g_exp2_table256[(256 / ETAB) * i];
    if (warp == EPI_WARPS) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(dst));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    ai_fence_before();
    __syncthreads();
    ai_fence_after();
    const uint32_t tmem = tmem_base_s;
    pdl_wait();                                            // barriers, TMEM and the table are set up while the slicers drain
    pdl_trigger();

    const int8_t* gK = Kd.Ks + (size_t)bh * T * (S * AI_KPLANE);
    const int8_t* gV = Kd.Vs + (size_t)bh * T * (S * AI_KPLANE);

    if (warp == EPI_WARPS + 1) {
        // ------------------------------------------------------------------ loader
        if (elect_one()) {
            const int8_t* gQ = Qd.Qs + ((size_t)bh * (Npad / AI_BM) + qt) * (S * AI_QPLANE);
            const unsigned sc_bytes = (unsigned)(Mpad * 4 + Mpad * 4 + ((T + 3) & ~3) * 4);
            mbar_expect_tx(&q_full, S * AI_QPLANE + sc_bytes);
#pragma unroll
            for (int s = 0; s < S; ++s) bulk_g2s(sQ + s * AI_QPLANE, gQ + (size_t)s * AI_QPLANE, AI_QPLANE, &q_full);
            bulk_g2s(s_kex, Kd.kexp + (size_t)bh * Mpad, Mpad * 4, &q_full);
            bulk_g2s(s_ksf, Kd.kscale_f + (size_t)bh * Mpad, Mpad * 4, &q_full);
            bulk_g2s(s_ktm, Kd.ktilemax + (size_t)bh * ((T + 3) & ~3), ((T + 3) & ~3) * 4, &q_full);
            // Ring units: in pass 1 a unit is a GROUP of up to P1G tiles (only the two leading K planes of each, 2 KB per
            // tile: one barrier round trip per tile would leave the ring, not the epilogue, as the bound of pass 1);
            // in pass 2 a unit is one tile (its S planes of K and of V^T).
            int u = 0;
            if (PASS1) {
                for (int g0 = 0; g0 < T; g0 += P1G, ++u) {
                    const int stage = u % AI_STAGES, nt = min(P1G, T - g0);
                    if (u >= AI_STAGES) mbar_wait(&kv_empty[stage], (unsigned)((u / AI_STAGES - 1) & 1));
                    mbar_expect_tx(&kv_full[stage], nt * 2 * AI_KPLANE);
                    for (int i = 0; i < nt; ++i)
                        bulk_g2s(sKV + stage * STAGE_BYTES + i * 2 * AI_KPLANE, gK + (size_t)(g0 + i) * (S * AI_KPLANE), 2 * AI_KPLANE, &kv_full[stage]);
                }
            }
            for (int jt = 0; jt < T; ++jt, ++u) {
                const int stage = u % AI_STAGES;
                if (u >= AI_STAGES) mbar_wait(&kv_empty[stage], (unsigned)((u / AI_STAGES - 1) & 1));
                mbar_expect_tx(&kv_full[stage], (LOGITS ? 1 : 2) * S * AI_KPLANE);
                bulk_g2s(sKV + stage * STAGE_BYTES, gK + (size_t)jt * (S * AI_KPLANE), S * AI_KPLANE, &kv_full[stage]);
                if (!LOGITS)
                    bulk_g2s(sKV + stage * STAGE_BYTES + S * AI_KPLANE, gV + (size_t)jt * (S * AI_KPLANE), S * AI_KPLANE, &kv_full[stage]);
            }
        }
    } else if (warp == EPI_WARPS) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const uint64_t qd0 = ai_desc(sQ), kd0 = ai_desc(sKV), pd0 = ai_desc(sP);
            mbar_wait(&q_full, 0);
            int u = 0;
            if (PASS1) {
                for (int g0 = 0; g0 < T; g0 += P1G, ++u) {
                    const int stage = u % AI_STAGES, nt = min(P1G, T - g0);
                    mbar_wait(&kv_full[stage], (unsigned)((u / AI_STAGES) & 1));
                    for (int i = 0; i < nt; ++i) {
                        const int jt = g0 + i, buf = jt & 3;
                        if (jt >= 4) mbar_wait(&s1_empty[buf], (unsigned)(((jt >> 2) - 1) & 1));
                        ai_fence_after();
                        const uint64_t kd = kd0 + (uint64_t)((stage * STAGE_BYTES + i * 2 * AI_KPLANE) >> 4);
                        const uint32_t d = tmem + (buf < 2 ? TM_P1_LO + buf * 64 : TM_P1_HI + (buf - 2) * 64);
                        // diagonal 0 = D0 G0, diagonal 1 = D0 G1 + D1 G0
                        ai_mma(d, qd0, kd, ai_idesc(64, true), false);
                        ai_mma(d + 32, qd0 + (uint64_t)(AI_QPLANE >> 4), kd, ai_idesc(32, true), true);
                        ai_commit(&s1_full[buf]);
                    }
                    ai_commit(&kv_empty[stage]);
                }
            }
            const int u0 = u;                                        // ring unit of pass-2 tile 0
            // Q K^T of tile jt into accumulator set jt % NSBUF (its (jt / NSBUF)-th use)
            auto qk_tile = [&](int jt) {
                const int uu = u0 + jt, stage = uu % AI_STAGES;
                mbar_wait(&kv_full[stage], (unsigned)((uu / AI_STAGES) & 1));
                const int sb = jt % NSBUF, use = jt / NSBUF;
                if (use >= 1) mbar_wait(&s_empty[sb], (unsigned)((use - 1) & 1));      // the epilogue has read the previous tile of this set
                else if (PASS1 && sb == NSBUF - 1) mbar_wait(&p1_done, 0);          // pass 1 borrowed the start of this set
                ai_fence_after();
                const uint64_t kd = kd0 + (uint64_t)((stage * STAGE_BYTES) >> 4);
#pragma unroll
                for (int s = 0; s < S; ++s)
                    ai_mma(tmem + TM_S + sb * (S * AI_BN) + s * AI_BN, qd0 + (uint64_t)((s * AI_QPLANE) >> 4), kd, ai_idesc((S - s) * AI_BN, true), s > 0);
                ai_commit(&s_full[sb]);
            };
            // The logits run NSBUF tiles ahead of P V: the epilogue asks for the accumulators of tile jt + 1 before it starts
            // the exponentials of tile jt, and P V of tile jt can only be issued once those exponentials are done.
            for (int t = 0; t < NSBUF && t < T; ++t) qk_tile(t);
            for (int jt = 0; jt < T; ++jt) {
                const int stage = (u0 + jt) % AI_STAGES;
                if (jt + NSBUF < T) qk_tile(jt + NSBUF);
                if (!LOGITS) {
                    const int buf = jt & 1;
                    mbar_wait(&p_full[buf], (unsigned)((jt >> 1) & 1));
                    ai_fence_after();
                    const uint64_t vd = kd0 + (uint64_t)((stage * STAGE_BYTES + S * AI_KPLANE) >> 4);
                    const uint64_t pd = pd0 + (uint64_t)((buf * SP * AI_PPLANE) >> 4);
#pragma unroll
                    for (int a = 0; a < SP; ++a)
                        ai_mma(tmem + TM_O + a * 32, pd + (uint64_t)((a * AI_PPLANE) >> 4), vd, ai_idesc((S - a) * 32, false), jt > 0 || a > 0);
                    ai_commit(&p_empty[buf]);
                }
                ai_commit(&kv_empty[stage]);
            }
            if (!LOGITS) ai_commit(&o_full);
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int quarter = warp & 3, cgi = warp >> 2, c0 = cgi * CW;   // columns c0 .. c0 + CW - 1 of every 32-column tile
        const int rloc = quarter * 32 + lane;                     // TMEM lane = row inside the query tile
        const int row = qt * AI_BM + rloc;
        const bool row_ok = row < N;
        const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
        if (!LOGITS && p.mp.Xs != nullptr && warp < 4) {
            // vscale = 2^(e_c - 13) of the 4 x 32 value channels of this pair: the largest exponent bounds every message entry
            const int hi = __double2hiint(Kd.vscale[(size_t)b * (HEADS * 32) + tid]);
            const int mx = __reduce_max_sync(0xffffffffu, hi);
            if (lane == 0) s_vexp[warp] = mx;                      // read after the epilogue barrier at the end
        }
        mbar_wait(&q_full, 0);                                     // key scales are in shared memory
        const double r_i = Qd.qscale[(size_t)bh * Npad + row];     // 2^(e_i - 12) / sqrt(32); 0 for padded rows
        // row factor of the recombined integer: even S: ai_recombine() returns 256 x the sum; CVT 4: 2^(16 (S-1)/2) x
        const double r_z = ((S & 1) ? r_i : r_i * 0.00390625) * ai_recombine_scale<S, CVT>();
        // XE: the row factor carries E / ln2 (ai_exp_fixed_z)
        const double r_zx = XE ? r_z * ((double)(XE ? XE : 1) * 1.4426950408889634) : r_z;
        const int r_zh = __double2hiint(r_zx), r_zl = __double2loint(r_zx);
        double c_i = 0.0, t_i = 0.0;
        int jl_i = 0;
        if (TOPK) {
            const long long grow = (long long)bh * N + row;
            c_i = row_ok ? p.tk.rmax[side][grow] : 0.0;                         // exact row maximum: p <= 1 with no slack
            t_i = row_ok ? p.tk.thr[side][grow] : INFINITY;                     // padded rows keep nothing
            jl_i = row_ok ? p.tk.jlast[side][grow] : -1;
        }
        if (PASS1) {
            // ---- pass 1: c_i >= max_j z_ij from the two leading diagonals
            float amax = -INFINITY, kmax = 0.f;
            // two tiles per trip: the TMEM loads of the second are in flight while the first is reduced (registers of an
            // asynchronous load are scoreboarded; tcgen05.wait::ld is only needed before the buffers are handed back)
            auto p1_load = [&](int jt, int (&a0)[CW], int (&a1)[CW]) {
                const int buf = jt & 3;
                mbar_wait(&s1_full[buf], (unsigned)((jt >> 2) & 1));
                ai_fence_after();
                const uint32_t col = buf < 2 ? TM_P1_LO + buf * 64 : TM_P1_HI + (buf - 2) * 64;
                ai_ldw(tlane + col + c0, a0);
                ai_ldw(tlane + col + 32 + c0, a1);
            };
            auto p1_use = [&](int jt, const int (&a0)[CW], const int (&a1)[CW]) {
                const float* kf = s_ksf + jt * AI_BN + c0;
                kmax = fmaxf(kmax, s_ktm[jt]);
                const int jbase = jt * AI_BN + c0;
                if (jbase + CW <= M) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) amax = fmaxf(amax, (float)(a0[j] * 256 + a1[j]) * kf[j]);   // units 2^(e_i - 20)
                } else {
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const float v = (float)(a0[j] * 256 + a1[j]) * kf[j];
                        if (jbase + j < M) amax = fmaxf(amax, v);
                    }
                }
            };
            for (int jt = 0; jt < T; jt += 2) {
                int a0[CW], a1[CW], b0[CW], b1[CW];
                const bool two = jt + 1 < T;
                p1_load(jt, a0, a1);
                if (two) p1_load(jt + 1, b0, b1);
                p1_use(jt, a0, a1);
                if (two) p1_use(jt + 1, b0, b1);
                ai_ld_wait();
                ai_fence_before();
                mbar_arrive(&s1_empty[jt & 3]);
                if (two) mbar_arrive(&s1_empty[(jt + 1) & 3]);
            }
            mbar_arrive(&p1_done);
            // |dropped digits| <= 24.1 * 2^(e_i + f_j - 12); fp32 rounding of the leading part <= 2^-23 relative
            const double lead = (double)amax * 0.00390625 * r_i;
            double c = lead + fabs(lead) * 4.76837158203125e-07 + 24.2 * (double)kmax * r_i;
            if (!(amax > -INFINITY)) c = -INFINITY;                               // this column group saw only padding columns
            s_xd[cgi * 128 + rloc] = c;
            epi_bar_sync<AI_EPI_THREADS>();
            c_i = c;
#pragma unroll
            for (int g = 0; g < NCG; ++g) c_i = fmax(c_i, s_xd[g * 128 + rloc]);
        }
        double CA = 0.0;
        if (XE) {
            // K = the bound in whole ln2 / E steps, rounded up (the 1e-6 covers the rounding of r_zx); MAGIC_N - K is exact
            const double K = ceil(fma(c_i, (double)(XE ? XE : 1) * 1.4426950408889634, 1e-6));
            CA = (6755399441055744.0 + (double)((XE ? XE : 1) * (8 * SP - 1))) - ((K > -1e12 && K < 1e12) ? K : 0.0);
        }
        uint32_t bsum[SP];                                         // byte sums of this thread's p^ per plane (DP4A), exact
#pragma unroll
        for (int a = 0; a < SP; ++a) bsum[a] = 0u;
        // The accumulator loads of tile jt + 1 are issued before the exponentials of tile jt and collected after them: the
        // TMEM latency hides behind ~250 instructions, and the registers they land in are dead in between.
        int acc[S][CW];
        double z[CW];
        auto s_load = [&](int jt) {
            const int sb = jt % NSBUF;
            mbar_wait(&s_full[sb], (unsigned)((jt / NSBUF) & 1));
            ai_fence_after();
#pragma unroll
            for (int dd = 0; dd < S; ++dd) ai_ldw(tlane + TM_S + sb * (S * AI_BN) + dd * AI_BN + c0, acc[dd]);
        };
        auto s_collect = [&](int jt) {
            ai_ld_wait();
            ai_fence_before();
            mbar_arrive(&s_empty[jt % NSBUF]);
#pragma unroll
            for (int j = 0; j < CW; ++j) z[j] = ai_recombine<S, CW, CVT>(acc, j);           // the integer Q K^T in units of the row / key scales
        };
        s_load(0);
        s_collect(0);
        for (int jt = 0; jt < T; ++jt) {
            const int jbase = jt * AI_BN + c0;
            // row factor x key scale 2^f_j: the key exponent is added into the exponent field of r_z (integer pipe)
            double rk[CW];
            {
                const int* kx = s_kex + jbase;
#pragma unroll
                for (int j = 0; j < CW; ++j) rk[j] = __hiloint2double(r_zh + kx[j], r_zl);
            }
            if (LOGITS) {
                // A thread owns 64 bytes of ITS row: stored from here a warp would touch 32 rows per instruction (9 B / clk / SM,
                // the kernel would be store-bound at 2.3 TB/s). The tile goes through shared memory instead and leaves as
                // 256-byte row segments, two rows per warp instruction.
                if (jt > 0) epi_bar_sync<AI_EPI_THREADS>();                      // the previous tile has been read out
#pragma unroll
                for (int j = 0; j < CW; ++j) s_lg[rloc * 33 + c0 + j] = __dmul_rn(z[j], rk[j]);
                if (jt + 1 < T) s_load(jt + 1);
                epi_bar_sync<AI_EPI_THREADS>();
                {
                    const int tile0 = jt * AI_BN, cc = (lane & 15) * 2;
                    const bool pair_ok = (M & 1) == 0;
#pragma unroll
                    for (int it = 0; it < 64 / EPI_WARPS; ++it) {
                        const int r = (warp * (64 / EPI_WARPS) + it) * 2 + (lane >> 4);
                        const int grow = qt * AI_BM + r;
                        const double v0 = s_lg[r * 33 + cc], v1 = s_lg[r * 33 + cc + 1];
                        if (grow < N) {
                            double* dst = p.Out[side] + ((long long)bh * N + grow) * (long long)M + tile0 + cc;
                            if (pair_ok && tile0 + cc + 1 < M) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                            else {
                                if (tile0 + cc < M) dst[0] = v0;
                                if (tile0 + cc + 1 < M) dst[1] = v1;
                            }
                        }
                    }
                }
                if (jt + 1 < T) s_collect(jt + 1);
                continue;
            }
            if (NSBUF == 2 && jt + 1 < T) s_load(jt + 1);            // one accumulator set: tile jt + 1 is still being multiplied
            uint32_t lo[CW], hi[SP > 4 ? CW : 1];
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                uint32_t h32;
                if (TOPK) {
                    // the scaled logit exactly as LOGITS mode stored it (one rounded product), compared with the row's
                    // threshold; ties at the threshold are kept up to column jl_i
                    const double zz = __dmul_rn(z[j], rk[j]);
                    const bool keep = zz > t_i || (zz == t_i && jbase + j <= jl_i);
                    ai_exp_fixed<SP>(zz - c_i, etab_s32, lo[j], h32);
                    if (!keep) { lo[j] = 0u; h32 = 0u; }
                } else if constexpr (XE != 0) {
                    ai_exp_fixed_z<SP <= 4 ? SP : 4, XE ? XE : 16>(z[j], rk[j], CA, etab_s32, lo[j]);
                    h32 = 0u;
                } else {
                    ai_exp_fixed<SP>(fma(z[j], rk[j], -c_i), etab_s32, lo[j], h32);    // p^ = rint(exp(z - c_i) 2^(8 SP - 1)), p <= 1
                }
                if (SP > 4) hi[j] = h32;
            }
            if (jbase + CW > M) {
#pragma unroll
                for (int j = 0; j < CW; ++j) if (jbase + j >= M) { lo[j] = 0u; if (SP > 4) hi[j] = 0u; }
            }
            // byte planes of P: plane a holds byte SP-1-a of every p^ (plane 0 = most significant); 4x4 byte transposes
            uint32_t w[SP][CW / 4];
#pragma unroll
            for (int g = 0; g < CW / 4; ++g) {
                const uint32_t a0 = __byte_perm(lo[4 * g], lo[4 * g + 1], 0x5140), a1 = __byte_perm(lo[4 * g], lo[4 * g + 1], 0x7362);
                const uint32_t b0 = __byte_perm(lo[4 * g + 2], lo[4 * g + 3], 0x5140), b1 = __byte_perm(lo[4 * g + 2], lo[4 * g + 3], 0x7362);
                w[SP - 1][g] = __byte_perm(a0, b0, 0x5410);
                w[SP - 2][g] = __byte_perm(a0, b0, 0x7632);
                w[SP - 3][g] = __byte_perm(a1, b1, 0x5410);
                if (SP >= 4) w[SP >= 4 ? SP - 4 : 0][g] = __byte_perm(a1, b1, 0x7632);
                if (SP > 4) {
                    const uint32_t c0h = __byte_perm(hi[4 * g], hi[4 * g + 1], 0x5140), d0h = __byte_perm(hi[4 * g + 2], hi[4 * g + 3], 0x5140);
                    w[SP - 5][g] = __byte_perm(c0h, d0h, 0x5410);
                    if (SP > 5) w[SP > 5 ? SP - 6 : 0][g] = __byte_perm(c0h, d0h, 0x7632);
                }
#pragma unroll
                for (int a = 0; a < SP; ++a) bsum[a] = __dp4a(w[a][g], 0x01010101u, bsum[a]);
            }
            const int buf = jt & 1;
            if (jt >= 2) mbar_wait(&p_empty[buf], (unsigned)(((jt >> 1) - 1) & 1));   // P V of tile jt - 2 has read this buffer
            uint8_t* pdst = sP + buf * (SP * AI_PPLANE) + canon32(rloc, c0 >> 4) + (c0 & 15);
#pragma unroll
            for (int a = 0; a < SP; ++a) {
                if constexpr (CW == 16) *reinterpret_cast<uint4*>(pdst + a * AI_PPLANE) = make_uint4(w[a][0], w[a][1], w[a][2], w[a][3]);
                else *reinterpret_cast<uint2*>(pdst + a * AI_PPLANE) = make_uint2(w[a][0], w[a][1]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy stores -> tensor core reads
            mbar_arrive(&p_full[buf]);
            if (jt + 1 < T) {
                if (NSBUF == 1) s_load(jt + 1);
                s_collect(jt + 1);
            }
        }
        if (!LOGITS) {
            unsigned long long rsum = 0ull;
#pragma unroll
            for (int a = 0; a < SP; ++a) rsum += (unsigned long long)bsum[a] << (8 * (SP - 1 - a));
            s_xu[cgi * 128 + rloc] = rsum;
            mbar_wait(&o_full, 0);
            ai_fence_after();
            epi_bar_sync<AI_EPI_THREADS>();
            unsigned long long tot = 0ull;
#pragma unroll
            for (int g = 0; g < NCG; ++g) tot += s_xu[g * 128 + rloc];
            const double inv = pow2i(8 * SP - 1) / (double)tot;                  // 1 / (sum p^ 2^-(8 SP - 1))
            const double* vs = Kd.vscale + (size_t)bh * 32 + c0;
            double outv[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) outv[j] = 0.0;
#pragma unroll
            for (int dd = S - 1; dd >= 0; --dd) {
                int o[CW];
                ai_ldw(tlane + TM_O + dd * 32 + c0, o);
                ai_ld_wait();
#pragma unroll
                for (int j = 0; j < CW; ++j) outv[j] = fma(outv[j], 0.00390625, int_to_f64<0>(o[j]));
            }
            if (p.mp.Xs != nullptr) {
                // The message row as the digit planes of the next GEMM's A operand (slice_rows_kernel's format: x = 2^e sum_s
                // d_s 2^(1-7s), |d_s| <= 64, d_s = q_s - 128 q_(s-1), q_s = rint(x 2^(6-e) 128^s) read out of the mantissa).
                // e = exponent bound of the pair's source values instead of the row maximum of the message: |message| <=
                // max |v| because the probabilities are a convex combination -- no second pass over the 128 columns (4 CTAs).
                const int e = (max(max(s_vexp[0], s_vexp[1]), max(s_vexp[2], s_vexp[3])) >> 20) - 1023 + 13;
                const long long r = p.mp.row0[side] + (long long)b * N + row;
                const int kcol = h * HDIM + c0;
                int8_t* dst = p.mp.Xs + ((size_t)(r >> 7) * p.mp.S) * (128 * 128) +
                              (((int)(r & 127) >> 3) * 1024 + (kcol >> 4) * 128 + ((int)(r & 7)) * 16 + (kcol & 15));
                double x[CW];
                int qp[CW];
#pragma unroll
                for (int j = 0; j < CW; ++j) { x[j] = outv[j] * vs[j] * inv; qp[j] = 0; }
#pragma unroll 1
                for (int s = 0; s < p.mp.S; ++s) {
                    const double cs = pow2i(6 - e + 7 * s);
                    uint32_t w[CW / 4];
#pragma unroll
                    for (int g = 0; g < CW / 4; ++g) w[g] = 0u;
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const int q = __double2loint(fma(x[j], cs, 6755399441055744.0));
                        const int d = q - (qp[j] << 7);
                        qp[j] = q;
                        w[j >> 2] |= ((uint32_t)d & 0xffu) << (8 * (j & 3));
                    }
                    if (row_ok) {
                        if constexpr (CW == 16) *reinterpret_cast<uint4*>(dst + (size_t)s * (128 * 128)) = make_uint4(w[0], w[1], w[2], w[3]);
                        else *reinterpret_cast<uint2*>(dst + (size_t)s * (128 * 128)) = make_uint2(w[0], w[1]);
                    }
                }
                if (row_ok && h == 0 && cgi == 0) p.mp.rowscale[r] = pow2i(e - 12);
            } else if (row_ok) {
                double* dst = p.Out[side] + ((long long)b * N + row) * p.ldo + h * HDIM + c0;
#pragma unroll
                for (int j = 0; j < CW; j += 2)
                    *reinterpret_cast<double2*>(dst + j) = make_double2(outv[j] * vs[j] * inv, outv[j + 1] * vs[j + 1] * inv);
            }
        }
    }
    ai_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

static size_t attn_i8_smem(int M, int S, int SP) {
    const int T = (M + AI_BN - 1) / AI_BN, Mpad = T * AI_BN;
    return (size_t)S * AI_QPLANE + (size_t)AI_STAGES * 2 * S * AI_KPLANE + 2 * SP * AI_PPLANE +
           (size_t)Mpad * 8 + (size_t)((T + 3) & ~3) * 4 + 2 * 2048 + 512 * 8 + 512 * 8 + 128 * 33 * 8;     // table + its alignment slack, LOGITS staging tile
}

// the planes of one query tile plus the per-key scales of ALL M sources must fit in shared memory (sized for S = 7, SP = 6 so
// that the answer does not depend on the precision setting)
bool attn_i8_supported(int N, int M) { return N > 0 && M > 0 && attn_i8_smem(M, 7, 6) <= 200 * 1024; }

template <int S>
static cudaError_t slice_t(const double* Qh, const double* Kh, const double* Vh, const AttnI8Side& o, int B, int nqb, int nkb, cudaStream_t st) {
    if (nqb + nkb > 0) {
        slice_qk_kernel<S><<<nqb + nkb, 128, 0, st>>>(Qh, Kh, o, B, nqb);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        count_launch();
    }
    if (Vh) {
        slice_v_kernel<S><<<B * HEADS, 32 * SV_WARPS, 0, st>>>(Vh, o);
        count_launch();
    }
    return cudaGetLastError();
}

// Qh / Kh / Vh: head-major float64 buffers of ONE side (any of them may be null: those planes are not cut)
cudaError_t launch_attn_i8_slice(const double* Qh, const double* Kh, const double* Vh, const AttnI8Side& o, int B, cudaStream_t st) {
    const int n = o.n;
    if (B <= 0 || n <= 0) return cudaSuccess;
    const int nqb = Qh ? B * HEADS * ((n + AI_BM - 1) / AI_BM) : 0;
    const int nkb = Kh ? B * HEADS * ((pad_to(n, AI_BN) + 127) / 128) : 0;
    switch (o.S) {
        case 4: return slice_t<4>(Qh, Kh, Vh, o, B, nqb, nkb, st);
        case 5: return slice_t<5>(Qh, Kh, Vh, o, B, nqb, nkb, st);
        case 6: return slice_t<6>(Qh, Kh, Vh, o, B, nqb, nkb, st);
        case 7: return slice_t<7>(Qh, Kh, Vh, o, B, nqb, nkb, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_attn_i8_slice_sides(const double* const* Qh, const double* const* Kh, const double* const* Vh, const AttnI8Side* o,
                                       int B, cudaStream_t st) {
    static_assert(SV_WARPS == 16, "a q/k CTA of the fused slicer holds four 128-row blocks");
    SliceSide sd[2];
    for (int s = 0; s < 2; ++s) {
        const int n = o[s].n;
        sd[s].Qh = Qh[s]; sd[s].Kh = Kh[s]; sd[s].Vh = Vh[s]; sd[s].o = o[s];
        sd[s].nqb = B * HEADS * ((n + AI_BM - 1) / AI_BM);
        sd[s].nkb = B * HEADS * ((pad_to(n, AI_BN) + 127) / 128);
        sd[s].qk_ctas = (sd[s].nqb + sd[s].nkb + 3) / 4;
        sd[s].v_ctas = B * HEADS;
    }
    if (B <= 0 || o[0].n <= 0 || o[1].n <= 0) return cudaSuccess;
    if (o[0].S != o[1].S) return cudaErrorInvalidValue;
    const int grid_qk = sd[0].nqb + sd[0].nkb + sd[1].nqb + sd[1].nkb, grid_v = sd[0].v_ctas + sd[1].v_ctas;
    const size_t smem_qk = 4 * 32 * SQ_PITCH;
    auto go = [&](auto kqk, auto kv) -> cudaError_t {
        cudaError_t r = cudaFuncSetAttribute(kqk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_qk);
        if (r != cudaSuccess) return r;
        r = launch_pdl(kqk, dim3(grid_qk), dim3(128), smem_qk, st, sd[0], sd[1]);
        if (r != cudaSuccess) return r;
        return launch_pdl(kv, dim3(grid_v), dim3(32 * SV_WARPS), 0, st, sd[0], sd[1]);
    };
    cudaError_t e;
    switch (o[0].S) {
        case 4: e = go(slice_qk_sides_kernel<4>, slice_v_sides_kernel<4>); break;
        case 5: e = go(slice_qk_sides_kernel<5>, slice_v_sides_kernel<5>); break;
        case 6: e = go(slice_qk_sides_kernel<6>, slice_v_sides_kernel<6>); break;
        case 7: e = go(slice_qk_sides_kernel<7>, slice_v_sides_kernel<7>); break;
        default: return cudaErrorInvalidValue;
    }
    if (e == cudaSuccess) count_launch(2);
    return e;
}

template <int S, int SP, int CVT>
static cudaError_t attn_i8_go(const AttnI8Params& p, dim3 grid, size_t smem, int mode, int cw, cudaStream_t st) {
    auto go = [&](auto kern, int threads) -> cudaError_t {
        cudaError_t r = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (r != cudaSuccess) return r;
        return launch_pdl(kern, grid, dim3(threads), smem, st, p);
    };
    if (mode == AI_MODE_LOGITS) return cw == 16 ? go(attn_i8_kernel<S, SP, AI_MODE_LOGITS, 16, CVT>, ai_threads(16)) : go(attn_i8_kernel<S, SP, AI_MODE_LOGITS, 8, CVT>, ai_threads(8));
    if (mode == AI_MODE_TOPK) return cw == 16 ? go(attn_i8_kernel<S, SP, AI_MODE_TOPK, 16, CVT>, ai_threads(16)) : go(attn_i8_kernel<S, SP, AI_MODE_TOPK, 8, CVT>, ai_threads(8));
    return cw == 16 ? go(attn_i8_kernel<S, SP, AI_MODE_FULL, 16, CVT>, ai_threads(16)) : go(attn_i8_kernel<S, SP, AI_MODE_FULL, 8, CVT>, ai_threads(8));
}

// q[s] / kv[s]: digit planes of the query side and of the source side of grid side s; Out[s]: messages (rows x ldo),
// or in AI_MODE_LOGITS the dense scaled logits (B,4,N,M) of that side. SP: byte planes of P (S = 4: 3 or 4, 5: 4, 6: 5, 7: 6).
cudaError_t launch_attn_i8(const AttnI8Side* q, const AttnI8Side* kv, double* const* Out, int B, int nsides, int ldo,
                           int mode, const AttnI8TopK* tk, int SP, cudaStream_t st, const AttnI8MsgPlanes* mp) {
    AttnI8Params p;
    if (mp != nullptr && mode != AI_MODE_LOGITS) {
        if (mp->Xs == nullptr || mp->rowscale == nullptr || mp->S < 1 || mp->S > 7) return cudaErrorInvalidValue;
        p.mp = *mp;
    } else {
        p.mp = AttnI8MsgPlanes{nullptr, nullptr, 0, {0, 0}};
    }
    if (mode == AI_MODE_TOPK) {
        if (tk == nullptr) return cudaErrorInvalidValue;
        p.tk = *tk;
        if (nsides < 2) { p.tk.thr[1] = tk->thr[0]; p.tk.jlast[1] = tk->jlast[0]; p.tk.rmax[1] = tk->rmax[0]; }
    } else {
        p.tk = AttnI8TopK{{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    }
    int nmax = 0, mmax = 0;
    const int S = q[0].S;
    for (int s = 0; s < 2; ++s) {
        const int t = s < nsides ? s : 0;
        p.q[s] = q[t]; p.kv[s] = kv[t]; p.Out[s] = Out[t];
        nmax = q[t].n > nmax ? q[t].n : nmax;
        mmax = kv[t].n > mmax ? kv[t].n : mmax;
        if (q[t].S != S || kv[t].S != S) return cudaErrorInvalidValue;
    }
    p.B = B; p.ldo = ldo;
    if (B <= 0 || nmax <= 0) return cudaSuccess;
    const size_t smem = attn_i8_smem(mmax, S, SP);
    dim3 grid((nmax + AI_BM - 1) / AI_BM, HEADS, nsides * B);
    // MDGAT_ATTN_CW=16|8 (read once): columns of a key tile per epilogue warp, i.e. 8 or 16 epilogue warps
    static const int cw = [] { const char* v = getenv("MDGAT_ATTN_CW"); return v && v[0] == '1' ? 16 : 8; }();
    // MDGAT_ATTN_CVT=0|1|2 (read once): int32 -> float64 conversion of the epilogue, see int_to_f64(); the sweep setting
    // (5, 4) is built in all three variants, the others with the default
    static const int cvt = [] { const char* v = getenv("MDGAT_ATTN_CVT"); return v && v[0] >= '0' && v[0] <= '6' ? v[0] - '0' : AI_CVT_DEFAULT; }();
    cudaError_t e;
    switch (S * 10 + SP) {
        case 43: e = attn_i8_go<4, 3, AI_CVT_DEFAULT>(p, grid, smem, mode, cw, st); break;
        case 44: e = attn_i8_go<4, 4, AI_CVT_DEFAULT>(p, grid, smem, mode, cw, st); break;
        case 54: e = cvt == 0 ? attn_i8_go<5, 4, 0>(p, grid, smem, mode, cw, st)
                   : cvt == 1 ? attn_i8_go<5, 4, 1>(p, grid, smem, mode, cw, st)
                   : cvt == 2 ? attn_i8_go<5, 4, 2>(p, grid, smem, mode, cw, st)
                   : cvt == 3 ? attn_i8_go<5, 4, 3>(p, grid, smem, mode, cw, st)
                   : cvt == 5 ? attn_i8_go<5, 4, 5>(p, grid, smem, mode, cw, st)
                   : cvt == 6 ? attn_i8_go<5, 4, 6>(p, grid, smem, mode, cw, st)
                              : attn_i8_go<5, 4, 4>(p, grid, smem, mode, cw, st); break;
        case 65: e = attn_i8_go<6, 5, AI_CVT_DEFAULT>(p, grid, smem, mode, cw, st); break;
        case 76: e = attn_i8_go<7, 6, AI_CVT_DEFAULT>(p, grid, smem, mode, cw, st); break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

}  // namespace mdgat
