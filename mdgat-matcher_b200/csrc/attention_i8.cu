// float64-faithful multi-head attention on the 5th-generation tensor cores: Q K^T and P V of
// attention() / dynamic_attention() (/root/reference/models/mdgat.py:190-210) as exact int8 slice
// products (tcgen05.mma kind::i8, int32 accumulators in TMEM, operands staged by TMA bulk copies).
//
// tcgen05 has no f64 MMA kind, and the DMMA flash kernel (attention_f64.cu) spends 64 FP64-pipe FLOPs per
// logit on Q K^T and 64 on P V. Here both contractions leave the FP64 pipe:
//
//  * slice_qk_kernel / slice_v_kernel write every head vector as balanced base-256 digits,
//        x = 2^(e-54) * sum_{s=0..6} D_s 256^(6-s),   D_s in [-128, 127],
//    e = exponent of the row maximum for q and k rows, of the column (channel) maximum over all source
//    keypoints for v. 7 digits = 55 bits. Planes are laid out in the UMMA canonical no-swizzle K-major
//    order, one contiguous block per tile, so one cp.async.bulk brings a tile in.
//  * Q K^T: the 28 digit products D_s G_t^T with s + t <= 6 are accumulated exactly in int32; products with
//    the same s + t share a TMEM column group ("diagonal"). Stacked-N issue: plane s of Q against planes
//    0..6-s of K in ONE MMA of N = (7-s)*32 written 32*s columns into the accumulator set. The epilogue
//    recombines the 7 diagonals in float64 (Horner, neighbouring diagonals merged exactly in int32 first).
//  * softmax needs exp(z - max) <= 1 before P can be cut into digits, and an integer accumulator cannot be
//    rescaled when a running maximum moves. So the row maximum comes first: PASS 1 multiplies only the top
//    two diagonals (3 digit products) and takes the row maximum in fp32; a rigorous bound on what the
//    dropped digits can add turns it into c_i >= max_j z_ij with c_i - max <~ 1-2 (1-3 bits of P).
//  * PASS 2: p = exp(z - c_i) in float64 (table exp, common.cuh), p^ = rint(p 2^47) read straight out of
//    the mantissa as 6 unsigned bytes = the 6 digit planes of P, written to shared memory in A-operand
//    order. P V: unsigned P digits x signed V digits, 27 products with a + t <= 6, accumulated over ALL
//    source keypoints in TMEM (no online rescaling), one Horner pass per query row at the end, divided by
//    the exact integer row sum of the p^.
//
// Per logit the FP64 pipe sees 4 int->double conversions, 4 FMA, the exponential and one FMA for p^ (20
// instructions) against 80 in the DMMA kernel; digit handling runs on the integer pipe.
// LOGITS mode (dynamic layers) stops after the Horner pass and stores the scaled logits for the exact
// top-k selection kernel.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int AI_BM = 128;                 // query rows per CTA = TMEM lanes
constexpr int AI_BN = 32;                  // source keypoints per tile
constexpr int AI_S = 7;                    // digits of q, k, v
constexpr int AI_SP = 6;                   // digits of P (48 bits)
constexpr int AI_QPLANE = AI_BM * 32;      // bytes of one Q digit plane of a query tile
constexpr int AI_KPLANE = AI_BN * 32;      // bytes of one K (or V^T) digit plane of a source tile
constexpr int AI_PPLANE = AI_BM * AI_BN;   // bytes of one P digit plane
constexpr int AI_STAGES = 3;               // K/V tile ring
constexpr int AI_STAGE_BYTES = 2 * AI_S * AI_KPLANE;
// epilogue organisation: CW = columns of a 32-column tile per warp (16: 8 epilogue warps, 8: 16 epilogue warps = 4 per SM
// sub-partition); the MMA and loader warps follow the epilogue warps
constexpr int ai_epi_threads(int cw) { return 128 * (32 / cw); }
constexpr int ai_threads(int cw) { return ai_epi_threads(cw) + 64; }
constexpr int AI_TM_S = 0;                 // TMEM columns: logits diagonals [0, 224)
constexpr int AI_TM_O = AI_S * AI_BN;      //               P V diagonals   [224, 448); pass 1 borrows [224, 352)
constexpr int AI_EXP_LIMIT = 60;           // |exponent| clamp of the digit scales (values beyond 2^60 are out of range)

size_t attn_i8_q_bytes(int B, int n) { return (size_t)B * HEADS * ((n + AI_BM - 1) / AI_BM) * AI_S * AI_QPLANE; }
size_t attn_i8_kv_bytes(int B, int n) { return (size_t)B * HEADS * ((n + AI_BN - 1) / AI_BN) * AI_S * AI_KPLANE; }
static int pad_to(int n, int a) { return (n + a - 1) / a * a; }
static int tiles_pad4(int n) { return (((n + AI_BN - 1) / AI_BN) + 3) & ~3; }     // ktilemax row stride: 16-byte bulk copies

size_t attn_i8_side_bytes(int B, int n) {
    // Q planes | K planes | V planes | qscale[B*4*npad128] | kscale_d[B*4*npad32] | vscale[B*4*32] | kscale_f | ktilemax
    const size_t rq = (size_t)B * HEADS * pad_to(n, AI_BM), rk = (size_t)B * HEADS * pad_to(n, AI_BN);
    size_t b = attn_i8_q_bytes(B, n) + 2 * attn_i8_kv_bytes(B, n);
    b += rq * 8 + rk * 8 + (size_t)B * HEADS * 32 * 8 + rk * 4 + (size_t)B * HEADS * tiles_pad4(n) * 4;
    return (b + 255) / 256 * 256;
}

AttnI8Side attn_i8_carve(void* base, int B, int n) {
    AttnI8Side s;
    const size_t rq = (size_t)B * HEADS * pad_to(n, AI_BM), rk = (size_t)B * HEADS * pad_to(n, AI_BN);
    unsigned char* p = reinterpret_cast<unsigned char*>(base);
    s.Qs = reinterpret_cast<int8_t*>(p); p += attn_i8_q_bytes(B, n);
    s.Ks = reinterpret_cast<int8_t*>(p); p += attn_i8_kv_bytes(B, n);
    s.Vs = reinterpret_cast<int8_t*>(p); p += attn_i8_kv_bytes(B, n);
    s.qscale = reinterpret_cast<double*>(p); p += rq * 8;
    s.kscale = reinterpret_cast<double*>(p); p += rk * 8;
    s.vscale = reinterpret_cast<double*>(p); p += (size_t)B * HEADS * 32 * 8;
    s.kscale_f = reinterpret_cast<float*>(p); p += rk * 4;
    s.ktilemax = reinterpret_cast<float*>(p);
    s.n = n;
    return s;
}

DEVINL double pow2i(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// canonical K-major core-matrix offset inside a plane whose rows hold 32 bytes of K: 8 rows x 16 bytes contiguous,
// the two 16-byte K halves 128 B apart (LBO), 8-row groups 256 B apart (SBO)
DEVINL int canon32(int r, int khalf) { return (r >> 3) * 256 + khalf * 128 + (r & 7) * 16; }

// 7 balanced base-256 digits of 16 values -> w[plane][4] (16 bytes per plane, plane 0 = most significant)
DEVINL void digits16(const double* x, double sc, uint32_t (&w)[AI_S][4]) {
#pragma unroll
    for (int s = 0; s < AI_S; ++s) { w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u; }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        long long I = __double2ll_rn(x[i] * sc);           // |I| <= 2^54
#pragma unroll
        for (int s = AI_S - 1; s >= 1; --s) {
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
// kscale = 2^f (double and float copies), ktilemax = largest kscale of a 32-row tile
// ---------------------------------------------------------------------------------------------------
// blk128 / t128: index of the 128-row block and the thread inside it (the stand-alone kernel maps them to blockIdx /
// threadIdx, the fused per-layer kernel packs four of them into a 512-thread CTA)
DEVINL void slice_qk_body(const double* __restrict__ Qh, const double* __restrict__ Kh, const AttnI8Side& o, int nqb, int blk128, int t128) {
    const bool isq = blk128 < nqb;
    const int n = o.n;
    const int npad = isq ? (n + AI_BM - 1) / AI_BM * AI_BM : (n + AI_BN - 1) / AI_BN * AI_BN;
    const int bpb = (npad + 127) / 128;                              // blocks per (b, h)
    const int blk = isq ? blk128 : blk128 - nqb;
    const int bh = blk / bpb;
    const int i = (blk - bh * bpb) * 128 + t128;                     // padded row inside (b, h)
    if (i >= npad) return;                                           // K only; whole warps (npad multiple of 32)
    const double* src = (isq ? Qh : Kh) + ((long long)bh * n + i) * LDH_QK;
    // the row is read twice (maximum, then 16 values at a time for the digits; the second read hits L1): 16 instead of
    // 32 live doubles keep the fused slicer at two 512-thread CTAs per SM
    double mx = 0.0;
    if (i < n) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) { const double2 v = *reinterpret_cast<const double2*>(src + c); mx = fmax(mx, fmax(fabs(v.x), fabs(v.y))); }
    }
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);                                     // |x| < 2^e
    e = max(-AI_EXP_LIMIT, min(AI_EXP_LIMIT, e));
    const double sc = pow2i(54 - e);
    int8_t* dst;
    int plane;
    if (isq) {
        dst = o.Qs + ((size_t)bh * (npad / AI_BM) + (i >> 7)) * (AI_S * AI_QPLANE) + canon32(i & 127, 0);
        plane = AI_QPLANE;
        o.qscale[(size_t)bh * npad + i] = i < n ? pow2i(e - 12) * 0.17677669529663688110 : 0.0;
    } else {
        dst = o.Ks + ((size_t)bh * (npad / AI_BN) + (i >> 5)) * (AI_S * AI_KPLANE) + canon32(i & 31, 0);
        plane = AI_KPLANE;
        const float kf = i < n ? __int_as_float((127 + e) << 23) : 0.f;
        o.kscale[(size_t)bh * npad + i] = i < n ? pow2i(e) : 0.0;
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
            for (int c = 0; c < 16; c += 2) { const double2 v = *reinterpret_cast<const double2*>(src + half * 16 + c); x[c] = v.x; x[c + 1] = v.y; }
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = 0.0;
        }
        uint32_t w[AI_S][4];
        digits16(x, sc, w);
#pragma unroll
        for (int s = 0; s < AI_S; ++s)
            *reinterpret_cast<uint4*>(dst + (size_t)s * plane + half * 128) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

__global__ void __launch_bounds__(128)
slice_qk_kernel(const double* __restrict__ Qh, const double* __restrict__ Kh, AttnI8Side o, int B, int nqb) {
    slice_qk_body(Qh, Kh, o, nqb, (int)blockIdx.x, (int)threadIdx.x);
}

// ---------------------------------------------------------------------------------------------------
// Digits of V, transposed: the P V product contracts over source keypoints, so the B operand is V^T
// (32 channel rows x 32 keypoints of K per tile) and all keypoints of a (b, h) share one exponent per
// channel. One CTA per (b, h); lane = channel. vscale[c] = 2^(e_c - 13): with p^ = p 2^47 and the digit
// weights, message = vscale * Horner(P V diagonals) / (sum_j p^_j 2^-47).
// ---------------------------------------------------------------------------------------------------
constexpr int SV_WARPS = 16;     // one CTA per (b, h) (128 CTAs at cfg2): 16 warps keep enough loads in flight per SM
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
    const double sc = pow2i(54 - e);
    // unit = 16 consecutive keypoints: 16 bytes per plane and channel
    for (int unit = warp; unit < npad / 16; unit += SV_WARPS) {
        const int j0 = unit * 16;
        double x[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) x[jj] = (j0 + jj) < n ? V[(long long)(j0 + jj) * LDH_V + lane] : 0.0;
        uint32_t w[AI_S][4];
        digits16(x, sc, w);
        int8_t* dst = o.Vs + ((size_t)bh * (npad / AI_BN) + (j0 >> 5)) * (AI_S * AI_KPLANE) + canon32(lane, (j0 >> 4) & 1);
#pragma unroll
        for (int s = 0; s < AI_S; ++s)
            *reinterpret_cast<uint4*>(dst + (size_t)s * AI_KPLANE) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

__global__ void __launch_bounds__(32 * SV_WARPS)
slice_v_kernel(const double* __restrict__ Vh, AttnI8Side o) { slice_v_body(Vh, o, (int)blockIdx.x); }

// Both sides of a layer in ONE launch (the forward path): 512-thread CTAs, [q/k blocks of side 0 | v blocks of side 0 |
// q/k blocks of side 1 | v blocks of side 1]; a q/k CTA holds four 128-row blocks.
struct SliceSide { const double *Qh, *Kh, *Vh; AttnI8Side o; int nqb, nkb, qk_ctas, v_ctas; };
__global__ void __launch_bounds__(32 * SV_WARPS, 2)
slice_sides_kernel(const __grid_constant__ SliceSide s0, const __grid_constant__ SliceSide s1) {
    int blk = blockIdx.x;
    const bool second = blk >= s0.qk_ctas + s0.v_ctas;
    const SliceSide& sd = second ? s1 : s0;
    if (second) blk -= s0.qk_ctas + s0.v_ctas;
    if (blk < sd.qk_ctas) {
        const int blk128 = blk * 4 + ((int)threadIdx.x >> 7);
        if (blk128 < sd.nqb + sd.nkb) slice_qk_body(sd.Qh, sd.Kh, sd.o, sd.nqb, blk128, (int)threadIdx.x & 127);
    } else {
        slice_v_body(sd.Vh, sd.o, blk - sd.qk_ctas);
    }
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
DEVINL double int_to_f64(int v) {                                                    // exact, integer ALU + one DADD
    return __hiloint2double(0x43380000 + (v >> 31), v) - 6755399441055744.0;
}
// instruction descriptor: D = s32, B = signed 8-bit, both operands K-major, M = 128; A signed or unsigned
DEVINL constexpr uint32_t ai_idesc(int n, bool a_signed) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(AI_BM >> 4) << 24);
}

struct AttnI8Params {
    AttnI8Side q[2];             // digit planes of the QUERY side of grid side s
    AttnI8Side kv[2];            // digit planes of its SOURCE side
    double* Out[2];              // messages (rows x ldo) or, LOGITS, dense (B,4,N,M) logits
    int B, ldo;
};

// One CTA = 128 query rows of one (side, b, h), warp specialised:
//   warp 9 lane 0   loader: TMA bulk copies of the Q planes, the key scales and the K / V^T tile ring
//   warp 8 lane 0   MMA issuer
//   warps 0..7      epilogue: thread = query row (TMEM lane = 32 * (warp % 4) + lane), warps w and w + 4 split the
//                   32 columns of a tile 16 / 16
// (CW = 16; with CW = 8 there are 16 epilogue warps, four per lane quarter with 8 columns each, then the MMA and loader warps)
template <bool LOGITS, int CW>
__global__ void __launch_bounds__(ai_threads(CW), 1) attn_i8_kernel(const __grid_constant__ AttnI8Params p) {
    constexpr int AI_EPI_THREADS = ai_epi_threads(CW), EPI_WARPS = AI_EPI_THREADS / 32, NCG = 32 / CW;
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

    int8_t* sQ = reinterpret_cast<int8_t*>(ai_smem);                              // [7][4096]
    int8_t* sKV = sQ + AI_S * AI_QPLANE;                                          // [stage][K 7 planes | V 7 planes]
    uint8_t* sP = reinterpret_cast<uint8_t*>(sKV + AI_STAGES * AI_STAGE_BYTES);   // [2][6][4096]
    double* s_ksd = reinterpret_cast<double*>(sP + 2 * AI_SP * AI_PPLANE);        // [Mpad]
    float* s_ksf = reinterpret_cast<float*>(s_ksd + Mpad);                        // [Mpad]
    float* s_ktm = s_ksf + Mpad;                                                  // [T] (padded to 4)
    double* etab = reinterpret_cast<double*>(s_ktm + ((T + 3) & ~3));             // [256] 2^(j/256)
    double* s_xd = etab + 256;                                                     // [4][128] row exchange between column groups
    unsigned long long* s_xu = reinterpret_cast<unsigned long long*>(s_xd + 512); // [4][128]

    __shared__ __align__(8) uint64_t q_full, kv_full[AI_STAGES], kv_empty[AI_STAGES], s1_full[2], s1_empty[2],
        s_full, s_empty, p_full[2], p_empty[2], o_full;
    __shared__ uint32_t tmem_base_s;

    if (tid == 0) {
        mbar_init(&q_full, 1);
#pragma unroll
        for (int i = 0; i < AI_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s1_full[i], 1); mbar_init(&s1_empty[i], AI_EPI_THREADS);
            mbar_init(&p_full[i], AI_EPI_THREADS); mbar_init(&p_empty[i], 1);
        }
        mbar_init(&s_full, 1); mbar_init(&s_empty, AI_EPI_THREADS); mbar_init(&o_full, 1);
        mbar_fence_init();
    }
    exp_table256_to_shared(etab);
    if (warp == EPI_WARPS) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(dst));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    ai_fence_before();
    __syncthreads();
    ai_fence_after();
    const uint32_t tmem = tmem_base_s;

    const int8_t* gK = Kd.Ks + (size_t)bh * T * (AI_S * AI_KPLANE);
    const int8_t* gV = Kd.Vs + (size_t)bh * T * (AI_S * AI_KPLANE);

    if (warp == EPI_WARPS + 1) {
        // ------------------------------------------------------------------ loader
        if (elect_one()) {
            const int8_t* gQ = Qd.Qs + ((size_t)bh * (Npad / AI_BM) + qt) * (AI_S * AI_QPLANE);
            const unsigned sc_bytes = (unsigned)(Mpad * 8 + Mpad * 4 + ((T + 3) & ~3) * 4);
            mbar_expect_tx(&q_full, AI_S * AI_QPLANE + sc_bytes);
#pragma unroll
            for (int s = 0; s < AI_S; ++s) bulk_g2s(sQ + s * AI_QPLANE, gQ + (size_t)s * AI_QPLANE, AI_QPLANE, &q_full);
            bulk_g2s(s_ksd, Kd.kscale + (size_t)bh * Mpad, Mpad * 8, &q_full);
            bulk_g2s(s_ksf, Kd.kscale_f + (size_t)bh * Mpad, Mpad * 4, &q_full);
            bulk_g2s(s_ktm, Kd.ktilemax + (size_t)bh * ((T + 3) & ~3), ((T + 3) & ~3) * 4, &q_full);
            int u = 0;
            if (!LOGITS) {
                for (int jt = 0; jt < T; ++jt, ++u) {               // pass 1: the two leading K planes
                    const int stage = u % AI_STAGES;
                    if (u >= AI_STAGES) mbar_wait(&kv_empty[stage], (unsigned)((u / AI_STAGES - 1) & 1));
                    mbar_expect_tx(&kv_full[stage], 2 * AI_KPLANE);
                    bulk_g2s(sKV + stage * AI_STAGE_BYTES, gK + (size_t)jt * (AI_S * AI_KPLANE), 2 * AI_KPLANE, &kv_full[stage]);
                }
            }
            for (int jt = 0; jt < T; ++jt, ++u) {
                const int stage = u % AI_STAGES;
                if (u >= AI_STAGES) mbar_wait(&kv_empty[stage], (unsigned)((u / AI_STAGES - 1) & 1));
                mbar_expect_tx(&kv_full[stage], (LOGITS ? 1 : 2) * AI_S * AI_KPLANE);
                bulk_g2s(sKV + stage * AI_STAGE_BYTES, gK + (size_t)jt * (AI_S * AI_KPLANE), AI_S * AI_KPLANE, &kv_full[stage]);
                if (!LOGITS)
                    bulk_g2s(sKV + stage * AI_STAGE_BYTES + AI_S * AI_KPLANE, gV + (size_t)jt * (AI_S * AI_KPLANE), AI_S * AI_KPLANE, &kv_full[stage]);
            }
        }
    } else if (warp == EPI_WARPS) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const uint64_t qd0 = ai_desc(sQ), kd0 = ai_desc(sKV), pd0 = ai_desc(sP);
            mbar_wait(&q_full, 0);
            int u = 0;
            if (!LOGITS) {
                for (int jt = 0; jt < T; ++jt, ++u) {
                    const int stage = u % AI_STAGES, buf = jt & 1;
                    mbar_wait(&kv_full[stage], (unsigned)((u / AI_STAGES) & 1));
                    if (jt >= 2) mbar_wait(&s1_empty[buf], (unsigned)(((jt >> 1) - 1) & 1));
                    ai_fence_after();
                    const uint64_t kd = kd0 + (uint64_t)((stage * AI_STAGE_BYTES) >> 4);
                    const uint32_t d = tmem + AI_TM_O + buf * 64;
                    // diagonal 0 = D0 G0, diagonal 1 = D0 G1 + D1 G0
                    ai_mma(d, qd0, kd, ai_idesc(64, true), false);
                    ai_mma(d + 32, qd0 + (uint64_t)(AI_QPLANE >> 4), kd, ai_idesc(32, true), true);
                    ai_commit(&s1_full[buf]);
                    ai_commit(&kv_empty[stage]);
                }
            }
            auto issue_qk = [&](int stage) {
                const uint64_t kd = kd0 + (uint64_t)((stage * AI_STAGE_BYTES) >> 4);
#pragma unroll
                for (int s = 0; s < AI_S; ++s)
                    ai_mma(tmem + AI_TM_S + s * AI_BN, qd0 + (uint64_t)((s * AI_QPLANE) >> 4), kd, ai_idesc((AI_S - s) * AI_BN, true), s > 0);
            };
            mbar_wait(&kv_full[u % AI_STAGES], (unsigned)((u / AI_STAGES) & 1));
            ai_fence_after();
            issue_qk(u % AI_STAGES);
            ai_commit(&s_full);
            for (int jt = 0; jt < T; ++jt, ++u) {
                const int stage = u % AI_STAGES;
                if (jt + 1 < T) {
                    const int nst = (u + 1) % AI_STAGES;
                    mbar_wait(&kv_full[nst], (unsigned)(((u + 1) / AI_STAGES) & 1));
                    mbar_wait(&s_empty, (unsigned)(jt & 1));          // logits of tile jt are in registers
                    ai_fence_after();
                    issue_qk(nst);
                    ai_commit(&s_full);
                }
                if (!LOGITS) {
                    const int buf = jt & 1;
                    mbar_wait(&p_full[buf], (unsigned)((jt >> 1) & 1));
                    ai_fence_after();
                    const uint64_t vd = kd0 + (uint64_t)((stage * AI_STAGE_BYTES + AI_S * AI_KPLANE) >> 4);
                    const uint64_t pd = pd0 + (uint64_t)((buf * AI_SP * AI_PPLANE) >> 4);
#pragma unroll
                    for (int a = 0; a < AI_SP; ++a)
                        ai_mma(tmem + AI_TM_O + a * 32, pd + (uint64_t)((a * AI_PPLANE) >> 4), vd, ai_idesc((AI_S - a) * 32, false), jt > 0 || a > 0);
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
        mbar_wait(&q_full, 0);                                     // key scales are in shared memory
        const double r_i = Qd.qscale[(size_t)bh * Npad + row];     // 2^(e_i - 12) / sqrt(32); 0 for padded rows
        double c_i = 0.0;
        if (!LOGITS) {
            // ---- pass 1: c_i >= max_j z_ij from the two leading diagonals
            float amax = -INFINITY, kmax = 0.f;
            for (int jt = 0; jt < T; ++jt) {
                const int buf = jt & 1;
                mbar_wait(&s1_full[buf], (unsigned)((jt >> 1) & 1));
                ai_fence_after();
                int a0[CW], a1[CW];
                ai_ldw(tlane + AI_TM_O + buf * 64 + c0, a0);
                ai_ldw(tlane + AI_TM_O + buf * 64 + 32 + c0, a1);
                ai_ld_wait();
                ai_fence_before();
                mbar_arrive(&s1_empty[buf]);
                const float* kf = s_ksf + jt * AI_BN + c0;
                kmax = fmaxf(kmax, s_ktm[jt]);
                const int jbase = jt * AI_BN + c0;
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                    const float v = (float)(a0[j] * 256 + a1[j]) * kf[j];        // units 2^(e_i - 20)
                    if (jbase + j < M) amax = fmaxf(amax, v);
                }
            }
            // |dropped digits| <= 24.1 * 2^(e_i + f_j - 12); fp32 rounding of the leading part <= 2^-23 relative
            const double lead = (double)amax * 0.00390625 * r_i;
            double c = lead + fabs(lead) * 4.76837158203125e-07 + 24.2 * (double)kmax * r_i;
            if (!(amax > -INFINITY)) c = -INFINITY;                               // this half saw only padding columns
            s_xd[cgi * 128 + rloc] = c;
            epi_bar_sync<AI_EPI_THREADS>();
            c_i = c;
#pragma unroll
            for (int g = 0; g < NCG; ++g) c_i = fmax(c_i, s_xd[g * 128 + rloc]);
        }
        unsigned long long rsum = 0ull;
        for (int jt = 0; jt < T; ++jt) {
            mbar_wait(&s_full, (unsigned)(jt & 1));
            ai_fence_after();
            int acc[AI_S][CW];
#pragma unroll
            for (int dd = 0; dd < AI_S; ++dd) ai_ldw(tlane + AI_TM_S + dd * AI_BN + c0, acc[dd]);
            ai_ld_wait();
            ai_fence_before();
            mbar_arrive(&s_empty);
            const int jbase = jt * AI_BN + c0;
            const double* ks = s_ksd + jbase;
            double z[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                // |acc_dd| <= 7 * 32 * 2^14 < 2^22: two neighbouring diagonals merge exactly in int32
                double hsum = int_to_f64(acc[5][j] * 256 + acc[6][j]);
                hsum = fma(hsum, 1.52587890625e-05, int_to_f64(acc[3][j] * 256 + acc[4][j]));
                hsum = fma(hsum, 1.52587890625e-05, int_to_f64(acc[1][j] * 256 + acc[2][j]));
                hsum = fma(hsum, 1.52587890625e-05, int_to_f64(acc[0][j]));
                z[j] = hsum * ks[j];                                             // exact: ks is a power of two
            }
            if (LOGITS) {
                if (row_ok) {
                    double* dst = p.Out[side] + ((long long)bh * N + row) * (long long)M + jbase;
                    if (jbase + CW <= M && (M & 1) == 0) {
#pragma unroll
                        for (int j = 0; j < CW; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(z[j] * r_i, z[j + 1] * r_i);
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; ++j) if (jbase + j < M) dst[j] = z[j] * r_i;
                    }
                }
                continue;
            }
            uint32_t lo[CW], hi[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const double pj = exp_neg_abs47(fma(z[j], r_i, -c_i), etab);     // <= 1
                const double pm = fma(pj, 140737488355328.0, 6755399441055744.0);   // p 2^47 rounded into the mantissa
                lo[j] = (uint32_t)__double2loint(pm);
                hi[j] = (uint32_t)__double2hiint(pm) & 0xffffu;                  // bits 32..47 of p^
            }
            if (jbase + CW > M) {
#pragma unroll
                for (int j = 0; j < CW; ++j) if (jbase + j >= M) { lo[j] = 0u; hi[j] = 0u; }
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) rsum += ((unsigned long long)hi[j] << 32) | lo[j];
            const int buf = jt & 1;
            if (jt >= 2) mbar_wait(&p_empty[buf], (unsigned)(((jt >> 1) - 1) & 1));   // P V of tile jt - 2 has read this buffer
            uint8_t* pdst = sP + buf * (AI_SP * AI_PPLANE) + canon32(rloc, c0 >> 4) + (c0 & 15);
#pragma unroll
            for (int a = 0; a < AI_SP; ++a) {
                const int byte = AI_SP - 1 - a;                                  // plane 0 = most significant byte
                uint32_t w[CW / 4];
#pragma unroll
                for (int g = 0; g < CW / 4; ++g) {
                    uint32_t x0, x1, x2, x3;
                    if (byte < 4) { x0 = lo[4 * g]; x1 = lo[4 * g + 1]; x2 = lo[4 * g + 2]; x3 = lo[4 * g + 3]; }
                    else { x0 = hi[4 * g]; x1 = hi[4 * g + 1]; x2 = hi[4 * g + 2]; x3 = hi[4 * g + 3]; }
                    const uint32_t sel = (uint32_t)(byte & 3) | ((uint32_t)((byte & 3) + 4) << 4);   // byte of x, byte of y
                    const uint32_t t01 = __byte_perm(x0, x1, sel), t23 = __byte_perm(x2, x3, sel);
                    w[g] = __byte_perm(t01, t23, 0x5410);
                }
                if constexpr (CW == 16) *reinterpret_cast<uint4*>(pdst + a * AI_PPLANE) = make_uint4(w[0], w[1], w[2], w[3]);
                else *reinterpret_cast<uint2*>(pdst + a * AI_PPLANE) = make_uint2(w[0], w[1]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy stores -> tensor core reads
            mbar_arrive(&p_full[buf]);
        }
        if (!LOGITS) {
            s_xu[cgi * 128 + rloc] = rsum;
            mbar_wait(&o_full, 0);
            ai_fence_after();
            epi_bar_sync<AI_EPI_THREADS>();
            unsigned long long tot = 0ull;
#pragma unroll
            for (int g = 0; g < NCG; ++g) tot += s_xu[g * 128 + rloc];
            const double inv = 140737488355328.0 / (double)tot;                 // 1 / (sum p^ 2^-47)
            const double* vs = Kd.vscale + (size_t)bh * 32 + c0;
            double outv[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) outv[j] = 0.0;
#pragma unroll
            for (int dd = AI_S - 1; dd >= 0; --dd) {
                int o[CW];
                ai_ldw(tlane + AI_TM_O + dd * 32 + c0, o);
                ai_ld_wait();
#pragma unroll
                for (int j = 0; j < CW; ++j) outv[j] = fma(outv[j], 0.00390625, int_to_f64(o[j]));
            }
            if (row_ok) {
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

static size_t attn_i8_smem(int M) {
    const int T = (M + AI_BN - 1) / AI_BN, Mpad = T * AI_BN;
    return (size_t)AI_S * AI_QPLANE + (size_t)AI_STAGES * AI_STAGE_BYTES + 2 * AI_SP * AI_PPLANE +
           (size_t)Mpad * 12 + (size_t)((T + 3) & ~3) * 4 + 256 * 8 + 512 * 8 + 512 * 8;
}

bool attn_i8_supported(int N, int M) { return N > 0 && M > 0 && attn_i8_smem(M) <= 200 * 1024; }

// Qh / Kh / Vh: head-major float64 buffers of ONE side (any of them may be null: those planes are not cut)
cudaError_t launch_attn_i8_slice(const double* Qh, const double* Kh, const double* Vh, const AttnI8Side& o, int B, cudaStream_t st) {
    const int n = o.n;
    if (B <= 0 || n <= 0) return cudaSuccess;
    const int nqb = Qh ? B * HEADS * ((n + AI_BM - 1) / AI_BM) : 0;
    const int nkb = Kh ? B * HEADS * ((pad_to(n, AI_BN) + 127) / 128) : 0;
    if (nqb + nkb > 0) {
        slice_qk_kernel<<<nqb + nkb, 128, 0, st>>>(Qh, Kh, o, B, nqb);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        count_launch();
    }
    if (Vh) {
        slice_v_kernel<<<B * HEADS, 32 * SV_WARPS, 0, st>>>(Vh, o);
        count_launch();
    }
    return cudaGetLastError();
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
    const int grid = sd[0].qk_ctas + sd[0].v_ctas + sd[1].qk_ctas + sd[1].v_ctas;
    slice_sides_kernel<<<grid, 32 * SV_WARPS, 0, st>>>(sd[0], sd[1]);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) count_launch();
    return e;
}

// q[s] / kv[s]: digit planes of the query side and of the source side of grid side s; Out[s]: messages (rows x ldo),
// or with logits_only the dense scaled logits (B,4,N,M) of that side
cudaError_t launch_attn_i8(const AttnI8Side* q, const AttnI8Side* kv, double* const* Out, int B, int nsides, int ldo,
                           bool logits_only, cudaStream_t st) {
    AttnI8Params p;
    int nmax = 0, mmax = 0;
    for (int s = 0; s < 2; ++s) {
        const int t = s < nsides ? s : 0;
        p.q[s] = q[t]; p.kv[s] = kv[t]; p.Out[s] = Out[t];
        nmax = q[t].n > nmax ? q[t].n : nmax;
        mmax = kv[t].n > mmax ? kv[t].n : mmax;
    }
    p.B = B; p.ldo = ldo;
    if (B <= 0 || nmax <= 0) return cudaSuccess;
    const size_t smem = attn_i8_smem(mmax);
    dim3 grid((nmax + AI_BM - 1) / AI_BM, HEADS, nsides * B);
    cudaError_t e;
    // MDGAT_ATTN_CW=16|8 (read once): columns of a key tile per epilogue warp, i.e. 8 or 16 epilogue warps
    static const int cw = [] { const char* v = getenv("MDGAT_ATTN_CW"); return v && v[0] == '1' ? 16 : 8; }();
    auto go = [&](auto kern, int threads) -> cudaError_t {
        cudaError_t r = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (r != cudaSuccess) return r;
        kern<<<grid, threads, smem, st>>>(p);
        return cudaSuccess;
    };
    if (logits_only) e = cw == 16 ? go(attn_i8_kernel<true, 16>, ai_threads(16)) : go(attn_i8_kernel<true, 8>, ai_threads(8));
    else e = cw == 16 ? go(attn_i8_kernel<false, 16>, ai_threads(16)) : go(attn_i8_kernel<false, 8>, ai_threads(8));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

}  // namespace mdgat
