// Multi-head attention of the MDGAT GNN layers in float64 on sm_100a.
//
//  * attn_full_kernel: softmax(q k^T / sqrt(32)) v  (attention(), /root/reference/models/
//    mdgat.py:190-194) flash-style: the (B,4,N,M) logits/prob tensors the reference
//    materialises (268 MB each at B=32, N=M=512) never leave the SM. QK^T and PV run on
//    DMMA.8x8x4; the softmax probabilities are fed to the PV product straight from the
//    accumulator registers (the C-fragment of m8n8k4 holds columns {2q, 2q+1} of lane quad
//    position q, so taking "k index q" = column 2q+e of the key tile makes the C fragment
//    an A fragment with no shuffle; V rows are read in the same permuted order).
//  * topk_softmax_pv_kernel: dynamic_attention() (mdgat.py:196-210): exact-k selection per
//    (b, h, query) row over dense logits produced by the batched GEMM, softmax over the kept
//    k, and a sparse P.V (only k of M value rows are touched). Ties at the k-th logit are
//    resolved towards the lowest index, so exactly k entries are kept (torch.topk semantics).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int A_BM = 64, A_BN = 64, A_THREADS = 128, A_STAGES = 2;
constexpr size_t A_SMEM = ((size_t)A_STAGES * A_BN * (LDH_QK + LDH_V) + 64) * sizeof(double);
// logits only: no V stages (the table keeps its place behind the K stages), so four or five CTAs fit an SM
constexpr size_t A_SMEM_LOGITS = ((size_t)A_STAGES * A_BN * LDH_QK + 64) * sizeof(double);

// LOGITS_ONLY = true: the same Q K^T pipeline, but the scaled logits are written to Out as a dense
// (B,4,N,M) tensor (ldo = M) for the exact top-k selection below; no softmax, V is not read.
template <bool LOGITS_ONLY>
__global__ void __launch_bounds__(A_THREADS, LOGITS_ONLY ? 4 : 3)
attn_full_kernel(AttnSides ps, int B, int ldo, double scale) {
    extern __shared__ __align__(16) double smem[];
    double* Ks = smem;                                   // [stage][A_BN][LDH_QK]
    double* Vs = smem + A_STAGES * A_BN * LDH_QK;        // [stage][A_BN][LDH_V] (absent with LOGITS_ONLY)
    double* etab = Vs + (LOGITS_ONLY ? 0 : A_STAGES * A_BN * LDH_V);         // [64] 2^(j/64)
    exp_table_to_shared(etab);
    pdl_wait();
    pdl_trigger();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // blockIdx.z = side * B + b: both sides of a GNN layer (they share the layer weights and are
    // independent, mdgat.py:270) run in ONE launch, which also packs the waves better
    const int side = blockIdx.z >= B ? 1 : 0;
    const int h = blockIdx.y, b = blockIdx.z - side * B;
    const int N = ps.N[side], M = ps.M[side];
    if ((int)blockIdx.x * A_BM >= N) return;
    const double* __restrict__ Q = ps.Q[side];
    const double* __restrict__ K = ps.K[side];
    const double* __restrict__ V = ps.V[side];
    double* __restrict__ Out = ps.Out[side];
    const long long bh = (long long)b * HEADS + h;
    const double* Qbh = Q + bh * N * LDH_QK;
    const double* Kbh = K + bh * M * LDH_QK;
    const double* Vbh = V + bh * M * LDH_V;
    const int row_base = blockIdx.x * A_BM + warp * 16;
    const int qr = lane >> 2, qc = lane & 3;

    // Q fragments stay in registers for the whole kernel: A[mt][ks] = Q[row][4*ks + qc]
    double qa[2][8];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int row = row_base + mt * 8 + qr;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) qa[mt][ks] = row < N ? Qbh[(long long)row * LDH_QK + ks * 4 + qc] : 0.0;
    }

    double o[2][4][2];
    double m_run[2], l_run[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        m_run[mt] = -INFINITY; l_run[mt] = 0.0;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) o[mt][dt][0] = o[mt][dt][1] = 0.0;
    }

    const int nchunks = (M + A_BN - 1) / A_BN;

    // K/V chunks are contiguous in the head-major buffers (padded rows included): ONE bulk copy
    // (TMA engine, cp.async.bulk -> SASS UBLKCP) per operand and chunk, issued by a single thread
    // and completed on an mbarrier; the SM issues no per-thread copy instructions.
    __shared__ __align__(8) uint64_t full_bar[A_STAGES];
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < A_STAGES; ++i) mbar_init(&full_bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto load_chunk = [&](int c, int buf) {
        const int j0 = c * A_BN;
        const int rows = min(A_BN, M - j0);
        double* kd = Ks + buf * A_BN * LDH_QK;
        double* vd = Vs + buf * A_BN * LDH_V;
        if (tid == 0) {
            const unsigned kbytes = (unsigned)(rows * LDH_QK * sizeof(double));
            const unsigned vbytes = LOGITS_ONLY ? 0u : (unsigned)(rows * LDH_V * sizeof(double));
            mbar_expect_tx(&full_bar[buf], kbytes + vbytes);
            bulk_g2s(kd, Kbh + (long long)j0 * LDH_QK, kbytes, &full_bar[buf]);
            if (!LOGITS_ONLY) bulk_g2s(vd, Vbh + (long long)j0 * LDH_V, vbytes, &full_bar[buf]);
        }
        if (rows < A_BN) {
            // ragged last chunk: rows the copy does not write must not hold stale NaN/Inf bit patterns
            // (their probabilities are exactly 0, and 0 * NaN would poison the message)
            for (int i = rows * LDH_QK + tid; i < A_BN * LDH_QK; i += A_THREADS) kd[i] = 0.0;
            if (!LOGITS_ONLY)
                for (int i = rows * LDH_V + tid; i < A_BN * LDH_V; i += A_THREADS) vd[i] = 0.0;
        }
    };

    load_chunk(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) load_chunk(c + 1, (c + 1) & 1);
        mbar_wait(&full_bar[c & 1], (unsigned)((c >> 1) & 1));
        if (c == 0) __syncthreads();        // zero-filled tail rows of a ragged first chunk
        const double* ks_ = Ks + (c & 1) * A_BN * LDH_QK;
        const double* vs_ = Vs + (c & 1) * A_BN * LDH_V;

        // ---- S = Q K^T for 16 rows x 64 columns per warp
        double s[2][8][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) s[mt][nt][0] = s[mt][nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const double bfrag = ks_[(nt * 8 + qr) * LDH_QK + ks * 4 + qc];
                dmma884(s[0][nt][0], s[0][nt][1], qa[0][ks], bfrag);
                dmma884(s[1][nt][0], s[1][nt][1], qa[1][ks], bfrag);
            }
        }
        const int j0 = c * A_BN;
        if (LOGITS_ONLY) {
            // scores = q.k / sqrt(d) (mdgat.py:201), row-major (b, h, n, m)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int row = row_base + mt * 8 + qr;
                if (row < N) {
                    double* srow = Out + (bh * N + row) * (long long)M + j0 + 2 * qc;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const int col = j0 + nt * 8 + 2 * qc;
                        const double y0 = s[mt][nt][0] * scale, y1 = s[mt][nt][1] * scale;
                        if (col + 1 < M && (M & 1) == 0) *reinterpret_cast<double2*>(srow + nt * 8) = make_double2(y0, y1);
                        else { if (col < M) srow[nt * 8] = y0; if (col + 1 < M) srow[nt * 8 + 1] = y1; }
                    }
                }
            }
            __syncthreads();
            continue;
        }
        // columns past M (last chunk only)
        if (j0 + A_BN > M) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (j0 + nt * 8 + 2 * qc + e >= M) { s[0][nt][e] = -INFINITY; s[1][nt][e] = -INFINITY; }
        }

        // ---- online softmax (raw dots; 1/sqrt(d) applied inside the exponent)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            double mx = s[mt][0][0];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { mx = fmax(mx, s[mt][nt][0]); mx = fmax(mx, s[mt][nt][1]); }
            mx = fmax(mx, shfl_xor_d(mx, 1));
            mx = fmax(mx, shfl_xor_d(mx, 2));
            const double m_new = fmax(m_run[mt], mx);
            const double alpha = exp_fast_neg((m_run[mt] - m_new) * scale, etab);
            m_run[mt] = m_new;
            double rs = 0.0;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[mt][nt][0] = exp_fast_neg((s[mt][nt][0] - m_new) * scale, etab);
                s[mt][nt][1] = exp_fast_neg((s[mt][nt][1] - m_new) * scale, etab);
                rs += s[mt][nt][0] + s[mt][nt][1];
            }
            l_run[mt] = l_run[mt] * alpha + rs;            // per-lane partial; quad-reduced at the end
#pragma unroll
            for (int dt = 0; dt < 4; ++dt) { o[mt][dt][0] *= alpha; o[mt][dt][1] *= alpha; }
        }

        // ---- O += P V ; k index qc of step (nt, e) is key column nt*8 + 2*qc + e
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double* vrow = vs_ + (nt * 8 + 2 * qc + e) * LDH_V + qr;
#pragma unroll
                for (int dt = 0; dt < 4; ++dt) {
                    const double bfrag = vrow[dt * 8];
                    dmma884(o[0][dt][0], o[0][dt][1], s[0][nt][e], bfrag);
                    dmma884(o[1][dt][0], o[1][dt][1], s[1][nt][e], bfrag);
                }
            }
        }
        __syncthreads();
    }

    if (LOGITS_ONLY) return;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        double l = l_run[mt];
        l += shfl_xor_d(l, 1);
        l += shfl_xor_d(l, 2);
        const int row = row_base + mt * 8 + qr;
        if (row < N) {
            double* orow = Out + ((long long)b * N + row) * ldo + h * HDIM + 2 * qc;
#pragma unroll
            for (int dt = 0; dt < 4; ++dt)
                *reinterpret_cast<double2*>(orow + dt * 8) = make_double2(o[mt][dt][0] / l, o[mt][dt][1] / l);
        }
    }
}

static cudaError_t launch_attn(const AttnSides& ps, int B, int nsides, int ldo, bool logits_only, cudaStream_t st) {
    int nmax = 0;
    for (int s = 0; s < nsides; ++s) nmax = ps.N[s] > nmax ? ps.N[s] : nmax;
    if (B <= 0 || nmax <= 0) return cudaSuccess;
    dim3 grid((nmax + A_BM - 1) / A_BM, HEADS, nsides * B);
    const double scale = 1.0 / sqrt((double)HDIM);
    cudaError_t e;
    if (logits_only) {
        if ((e = cudaFuncSetAttribute(attn_full_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A_SMEM_LOGITS)) != cudaSuccess) return e;
        if ((e = launch_pdl(attn_full_kernel<true>, grid, dim3(A_THREADS), A_SMEM_LOGITS, st, ps, B, ldo, scale)) != cudaSuccess) return e;
    } else {
        if ((e = cudaFuncSetAttribute(attn_full_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A_SMEM)) != cudaSuccess) return e;
        if ((e = launch_pdl(attn_full_kernel<false>, grid, dim3(A_THREADS), A_SMEM, st, ps, B, ldo, scale)) != cudaSuccess) return e;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_attention_full(const AttnSides& ps, int B, int nsides, int ldo, cudaStream_t st) {
    return launch_attn(ps, B, nsides, ldo, false, st);
}

// Out[side] receives the dense scaled logits (B,4,N,M) of that side
cudaError_t launch_attention_logits(const AttnSides& ps, int B, int nsides, cudaStream_t st) {
    return launch_attn(ps, B, nsides, 0, true, st);
}

// ------------------------------------------------------------------------------------------
// Exact top-k + softmax + sparse PV (dynamic_attention(), mdgat.py:196-210). One warp per (b, h, query) row; lane l
// holds logits j = l + 32 v.
//
// Selection of exactly k entries, ties at the k-th value towards the lowest index:
//   1. bounds lo_b <= z_j <= hi_b of the row from the HIGH WORDS of the doubles alone (order-preserving 32-bit keys,
//      IMNMX + one REDUX each: no FP64 compare, no FP64 shuffle);
//   2. every logit is mapped to one of NB = 16 VPT equal-width bins by ONE fma whose result is read out of the mantissa
//      (u = rint(256 t), t = (z - lo_b) (NB-1)/(hi_b - lo_b) + 1/4, bin = u >> 8). The map is monotone, so an entry in a
//      higher bin is strictly larger than any entry in a lower one. A per-warp histogram in shared memory (one atomic
//      add per logit) and a suffix scan (NB / 32 bins per lane) give the bin b* that holds the k-th largest value and
//      the number of entries above it;
//   3. entries above b* are kept (lane counts, one warp scan, predicated stores: deterministic order); the handful of
//      entries INSIDE b* (M / NB x a density factor: 2-6 at M = 512) go to a list and are ranked exactly against each
//      other -- (value descending, column ascending) -- the best k - above of them are kept.
// ~250 instructions per row against ~1000 for the probe-and-count search this replaces (96 dependent FP64 compare /
// reduce rounds in the worst case). Rows the fast path cannot take -- all logits equal, a boundary bin with more than
// TK_LMAX entries (many exactly tied logits: duplicated keypoints, load_data.py:198-201), k >= M, non-finite bounds --
// go through topk_select_exact(): most-significant-bit-first search on the order-preserving 64-bit image.
// ------------------------------------------------------------------------------------------
DEVINL unsigned long long order_key(double x) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    return (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
}
DEVINL int order_key32(int hi) { return hi ^ ((hi >> 31) & 0x7fffffff); }      // signed order of the high words; an involution

constexpr int TK_WARPS = 8;
constexpr int TK_LMAX = 64;                  // capacity of the boundary-bin list
constexpr double TK_MAGIC = 26388279066624.0;  // 1.5 * 2^44: the low mantissa word of MAGIC + t is rint(256 t) for 0 <= t < 2^24

// A kept entry: logit (later probability) and the BYTE offset of its value row inside the head's value matrix
// (column * LDH_V * 8). 16-byte slots, but the P.V loop reads the two fields with an 8- and a 4-byte broadcast load
// (one shared-memory wavefront each; a 16-byte load of the struct costs four).
struct __align__(16) KeptEntry { double p; int voff; int pad; };
constexpr int TK_VROW = LDH_V * 8;           // bytes of a value row

// The slow, always-correct selection (cold path): reloads the row, finds the k-th largest order key bit by bit, keeps
// everything above it and the lowest-index ties. Writes exactly min(k, M) entries in column order.
template <int VPT>
__device__ __noinline__ void topk_select_exact(const double* __restrict__ srow, int M, int topk, int lane, KeptEntry* kept) {
    unsigned long long key[VPT];                 // 0 for padding: below every real key (real keys have a bit set)
    unsigned long long kmax = 0ull, kmin = ~0ull;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const int j = lane + 32 * v;
        key[v] = j < M ? order_key(srow[j]) : 0ull;
        if (j < M) { kmax = key[v] > kmax ? key[v] : kmax; kmin = key[v] < kmin ? key[v] : kmin; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmax, o), b = __shfl_xor_sync(0xffffffffu, kmin, o);
        kmax = a > kmax ? a : kmax; kmin = b < kmin ? b : kmin;
    }
    unsigned long long prefix = kmin;            // k >= M: everything real is kept
    bool found = topk >= M;
    if (!found) {
        const unsigned long long diff = kmax ^ kmin;
        prefix = kmax;
        if (diff != 0ull) {
            const int top = 63 - __clzll((long long)diff);
            prefix = (top == 63) ? 0ull : (kmax >> (top + 1)) << (top + 1);
            for (int bit = top; bit >= 0; --bit) {
                const unsigned long long cand = prefix | (1ull << bit);
                int c = 0;
#pragma unroll
                for (int v = 0; v < VPT; ++v) c += (key[v] >= cand) ? 1 : 0;
                c = __reduce_add_sync(0xffffffffu, c);
                if (c >= topk) {
                    prefix = cand;
                    if (c == topk) { found = true; break; }
                }
            }
        }
    }
    int gt = 0;
    if (!found) {
#pragma unroll
        for (int v = 0; v < VPT; ++v) gt += (key[v] > prefix) ? 1 : 0;
        gt = __reduce_add_sync(0xffffffffu, gt);
    }
    const int need = topk - gt;                  // tied entries to keep, lowest index first
    int seen = 0, base = 0;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const unsigned long long kv = key[v];
        bool take;
        if (found) {
            take = kv >= prefix && kv != 0ull;
        } else {
            const bool eq = kv == prefix;
            const unsigned em = __ballot_sync(0xffffffffu, eq);
            take = (kv > prefix) || (eq && (seen + __popc(em & ((1u << lane) - 1u))) < need);
            seen += __popc(em);
        }
        const unsigned tm = __ballot_sync(0xffffffffu, take);
        if (take) {
            const unsigned long long bits = (kv >> 63) ? (kv ^ 0x8000000000000000ull) : ~kv;
            KeptEntry e; e.p = __longlong_as_double((long long)bits); e.voff = (lane + 32 * v) * TK_VROW; e.pad = 0;
            kept[base + __popc(tm & ((1u << lane) - 1u))] = e;
        }
        base += __popc(tm);
    }
    __syncwarp();
}

// The fast selection described above. s[]: the row (padding = -inf; FULLROW: M == 32 VPT, no padding). Returns false
// when the row must take the exact path (nothing useful has been written then). The histogram aliases the start of
// kept[] (NB counters; kept is aligned to 4 NB bytes so that a counter address is base | offset), alist is the warp's
// boundary-bin list. hi_b (>= every logit, within 2^-20 relative of the maximum) is the softmax shift.
template <int VPT, bool FULLROW>
DEVINL bool topk_select_fast(const double (&s)[VPT], int M, int topk, int lane, KeptEntry* kept, KeptEntry* alist, double& hi_b) {
    constexpr int NB = 16 * VPT, BPL = NB / 32;
    constexpr bool KEEP_U = false;               // the 32-bit bin keys are recomputed (one fma) rather than kept: registers
    unsigned* hist = reinterpret_cast<unsigned*>(kept);
    const uint32_t kept_s = (uint32_t)__cvta_generic_to_shared(kept), alist_s = (uint32_t)__cvta_generic_to_shared(alist);
#pragma unroll
    for (int q = 0; q < BPL / 4; ++q) reinterpret_cast<uint4*>(hist)[lane + 32 * q] = make_uint4(0u, 0u, 0u, 0u);
    int kmax = (int)0x80000000, kmin = 0x7fffffff;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const int key = order_key32(__double2hiint(s[v]));
        kmax = max(kmax, key);
        kmin = min(kmin, (FULLROW || (lane + 32 * v) < M) ? key : 0x7fffffff);
    }
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    const int hmax = order_key32(kmax), hmin = order_key32(kmin);
    hi_b = __hiloint2double(hmax, hmax < 0 ? 0 : -1);
    const double lo_b = __hiloint2double(hmin, hmin < 0 ? -1 : 0);
    const double range = hi_b - lo_b;
    // a little less than (NB - 1) / range, in float32 with the reciprocal rounded down (any smaller positive scale keeps
    // the bins inside [0, NB)): a handful of instructions instead of a float64 division
    const double scale = (double)(__frcp_rd(__double2float_ru(range)) * ((float)(NB - 1) * 0.99999f));
    // |lo_b| scale < 2^40 keeps MAGIC + 1/4 - lo_b scale on the 2^-8 grid of MAGIC (NaN / Inf bounds fail the tests too)
    if (!(range > 1.0e-30 && range < 1.0e30 && fmax(fabs(lo_b), fabs(hi_b)) * scale < 1.0e12) || topk >= M) return false;
    const double off = (TK_MAGIC + 0.25) - lo_b * scale;
    auto ukey = [&](int v) -> unsigned { return (unsigned)__double2loint(fma(s[v], scale, off)); };   // -inf -> 0
    unsigned u[KEEP_U ? VPT : 1];
    __syncwarp();
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const unsigned uv = ukey(v);
        if (KEEP_U) u[v] = uv;
        // counter address = base | 4 * bin (the base is aligned to the size of the histogram)
        const uint32_t addr = kept_s | ((uv >> 6) & (unsigned)(4 * NB - 4));
        asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory");
    }
    __syncwarp();
    // suffix scan from the top bin: lane l owns bins [BPL l, BPL (l + 1))
    unsigned h[BPL];
#pragma unroll
    for (int q = 0; q < BPL / 4; ++q) {
        const uint4 t = reinterpret_cast<const uint4*>(hist)[lane * (BPL / 4) + q];
        h[4 * q] = t.x; h[4 * q + 1] = t.y; h[4 * q + 2] = t.z; h[4 * q + 3] = t.w;
    }
    int sfx[BPL + 1];                                        // sfx[b] = entries of this lane's bins b .. BPL-1
    sfx[BPL] = 0;
#pragma unroll
    for (int b = BPL - 1; b >= 0; --b) sfx[b] = sfx[b + 1] + (int)h[b];
    const int tot = sfx[0];
    int suf = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_down_sync(0xffffffffu, suf, o);
        if (lane + o < 32) suf += t;
    }
    const int above_lane = suf - tot;
    const bool mine = above_lane < topk && topk <= suf;      // exactly one lane (1 <= k <= M <= total count)
    // inside the owning lane: b* = number of bins b >= 1 whose suffix (from b up) still reaches k
    // (walking down from the top bin: the last suffix that does not reach k is the count above b*)
    int nb = 0, abv = above_lane;
#pragma unroll
    for (int b = BPL - 1; b >= 1; --b) {
        const int c = above_lane + sfx[b];
        nb += (c >= topk) ? 1 : 0;
        abv = (c >= topk) ? abv : c;
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
    const int bstar = __shfl_sync(0xffffffffu, lane * BPL + nb, src);
    const int above = __shfl_sync(0xffffffffu, abv, src);
    // padding has u = 0 and every real entry u >= 62 (t >= 1/4 - rounding), so a lower limit of at least 1 keeps the
    // padding out of the boundary list without a column test
    const unsigned t_hi = (unsigned)(bstar + 1) << 8, t_lo = max((unsigned)bstar << 8, 1u);
    // One counting pass (entries above / inside the boundary bin per lane, packed into one word), one warp scan, one
    // write pass: the entries above go to kept[] in (lane, v) order, the few inside to the boundary list. Both passes
    // are straight predicated code -- written as PTX so that they stay free of per-element branches. The histogram is
    // dead (every lane passed the shuffles above after reading its bins): kept[] may be overwritten.
    int cnt = 0;                                             // low half: above, high half: inside
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const unsigned uv = KEEP_U ? u[v] : ukey(v);
        asm("{\n.reg .pred p, q;\nsetp.ge.u32 p, %1, %2;\nsetp.ge.and.u32 q, %1, %3, !p;\n@p add.s32 %0, %0, 1;\n@q add.s32 %0, %0, 65536;\n}"
            : "+r"(cnt) : "r"(uv), "r"(t_hi), "r"(t_lo));
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const int L = __shfl_sync(0xffffffffu, inc, 31) >> 16;
    if (L > TK_LMAX) return false;
    uint32_t kp = kept_s + (uint32_t)((inc - cnt) & 0xffff) * 16u, ap = alist_s + (uint32_t)((inc - cnt) >> 16) * 16u;
    const int voff0 = lane * TK_VROW;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const unsigned uv = KEEP_U ? u[v] : ukey(v);
        // one 16-byte store per kept entry: {logit, row offset}
        asm volatile("{\n.reg .pred p, q;\n.reg .b64 t, x;\nsetp.ge.u32 p, %2, %3;\nsetp.ge.and.u32 q, %2, %4, !p;\n"
                     "cvt.u64.u32 t, %6;\nmov.b64 x, %5;\n"
                     "@p st.shared.v2.b64 [%0], {x, t};\n@p add.u32 %0, %0, 16;\n"
                     "@q st.shared.v2.b64 [%1], {x, t};\n@q add.u32 %1, %1, 16;\n}"
                     : "+r"(kp), "+r"(ap) : "r"(uv), "r"(t_hi), "r"(t_lo), "d"(s[v]), "r"(voff0 + v * (32 * TK_VROW)) : "memory");
    }
    __syncwarp();
    // exact rank inside the boundary bin: (value descending, column ascending) is a strict total order, so the ranks
    // are a permutation of 0 .. L-1 and the entries of rank < need land on distinct slots
    const int need = topk - above;
    for (int t = lane; t < L; t += 32) {
        const KeptEntry et = alist[t];
        int rank = 0;
        for (int q = 0; q < L; ++q) {
            const KeptEntry eq = alist[q];
            rank += (eq.p > et.p || (eq.p == et.p && eq.voff < et.voff)) ? 1 : 0;
        }
        if (rank < need) kept[above + rank] = et;
    }
    __syncwarp();
    return true;
}

// softmax numerators of the kept entries of a row (mdgat.py:206-207), shifted by mx >= every kept logit; returns their sum
DEVINL double topk_softmax_kept(KeptEntry* kept, int nk, double mx, const double* etab, int lane) {
    double sum = 0.0;
    for (int t = lane; t < nk; t += 32) {
        const double e = exp_fast_neg(kept[t].p - mx, etab);
        kept[t].p = e;
        sum += e;
    }
    sum = warp_sum_d(sum);
    __syncwarp();
    return sum;
}

// Sparse P.V of one query row and the store of its 32 message channels (dst: channel 0 of this head in the row).
// Lane roles: quarter-warp qw takes the kept entries t = qw (mod 4); lane cl of it the channels 2cl, 2cl+1, 16+2cl,
// 17+2cl -- its two 16-byte reads of a value row fall on two contiguous 128-byte runs per quarter-warp. Per entry one
// 8-byte and one 4-byte broadcast read (probability, row offset) and 256 bytes of the value row; four entries per trip
// of the warp, two trips in flight. kq_s: shared address of kept[qw]; vq_s / vq_g: value matrix + 16 cl bytes.
template <bool SMEM_V>
DEVINL void topk_pv_store(uint32_t kq_s, uint32_t vq_s, const char* vq_g, int nk, double sum, int lane, double* dst) {
    const int qw = lane >> 3, cl = lane & 7;
    auto entry = [&](int n, double& pr, double2& x0, double2& x1) {
        int voff;
        asm volatile("ld.shared.f64 %0, [%2];\n ld.shared.b32 %1, [%2+8];" : "=d"(pr), "=r"(voff) : "r"(kq_s + (uint32_t)n * 64u));
        if (SMEM_V) {
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%4];\n ld.shared.v2.f64 {%2, %3}, [%4+128];"
                         : "=d"(x0.x), "=d"(x0.y), "=d"(x1.x), "=d"(x1.y) : "r"(vq_s + (uint32_t)voff));
        } else {
            x0 = __ldg(reinterpret_cast<const double2*>(vq_g + voff));
            x1 = __ldg(reinterpret_cast<const double2*>(vq_g + voff + 128));
        }
    };
    double2 a0 = make_double2(0.0, 0.0), a1 = a0, b0 = a0, b1 = a0;
    const int nq = (nk - qw + 3) >> 2;                   // entries of this quarter-warp
    int n = 0;
#pragma unroll 2
    for (; n + 1 < nq; n += 2) {
        double p0, p1; double2 x00, x01, x10, x11;
        entry(n, p0, x00, x01);
        entry(n + 1, p1, x10, x11);
        a0.x = fma(p0, x00.x, a0.x); a0.y = fma(p0, x00.y, a0.y); a1.x = fma(p0, x01.x, a1.x); a1.y = fma(p0, x01.y, a1.y);
        b0.x = fma(p1, x10.x, b0.x); b0.y = fma(p1, x10.y, b0.y); b1.x = fma(p1, x11.x, b1.x); b1.y = fma(p1, x11.y, b1.y);
    }
    if (n < nq) {
        double p0; double2 x00, x01;
        entry(n, p0, x00, x01);
        a0.x = fma(p0, x00.x, a0.x); a0.y = fma(p0, x00.y, a0.y); a1.x = fma(p0, x01.x, a1.x); a1.y = fma(p0, x01.y, a1.y);
    }
    a0.x += b0.x; a0.y += b0.y; a1.x += b1.x; a1.y += b1.y;
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
        a0.x += shfl_xor_d(a0.x, o); a0.y += shfl_xor_d(a0.y, o);
        a1.x += shfl_xor_d(a1.x, o); a1.y += shfl_xor_d(a1.y, o);
    }
    if (qw == 0) {
        const double inv = 1.0 / sum;
        *reinterpret_cast<double2*>(dst + 2 * cl) = make_double2(a0.x * inv, a0.y * inv);
        *reinterpret_cast<double2*>(dst + 16 + 2 * cl) = make_double2(a1.x * inv, a1.y * inv);
    }
}

// VPT = values per lane (M <= 32*VPT). Dynamic shared memory (aligned to 4 NB bytes): per warp max(topk, 4 VPT)
// KeptEntry rounded up to a multiple of 4 VPT (its first 64 VPT bytes double as the selection histogram), per warp
// TK_LMAX boundary-bin entries, then the 64-entry exp table.
//
// SMEM_V = false: one warp per row, 8 rows per CTA; the sparse P.V gathers its k value rows from global memory
//                 (L2): 32 KB per query row, 2.1 GB per side at cfg2 -- the L2 read bandwidth is the bound.
// SMEM_V = true : a CTA (24 warps) owns 1 / nsplit of the query rows of one (side, b, h) (grid = (b h, nsplit, side):
//                 nsplit is chosen so that the CTA count fills whole waves of the 148 SMs); the whole value matrix
//                 of the head (M x 34 doubles, 136 KB at M = 512) is brought in by ONE TMA bulk copy and every warp
//                 walks query rows i = r0 + warp, r0 + warp + 24, .. gathering from shared memory. The logits row of
//                 the next query is requested before the P.V of the current one.
constexpr int TKS_WARPS = 20;
struct TopkSides { const double* S[2]; const double* V[2]; double* Out[2]; int N[2], M[2]; };
DEVINL constexpr int tk_kept_entries(int topk, int vpt) { return (topk + 4 * vpt - 1) / (4 * vpt) * (4 * vpt); }

template <int VPT, bool SMEM_V, bool FULLROW>
__global__ void __launch_bounds__(SMEM_V ? 32 * TKS_WARPS : 32 * TK_WARPS, (SMEM_V || VPT != 16) ? 1 : 3)
topk_softmax_pv_kernel(const __grid_constant__ TopkSides ps, int ldo, int topk, int nbh) {
    extern __shared__ __align__(16) unsigned char tk_smem_raw[];
    constexpr int WARPS = SMEM_V ? TKS_WARPS : TK_WARPS;
    constexpr uint32_t HB = 64 * VPT;                                      // bytes of the histogram = alignment of a kept list
    const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(tk_smem_raw);
    unsigned char* tk_smem = tk_smem_raw + (((raw_s + HB - 1) & ~(HB - 1)) - raw_s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int side = blockIdx.z;
    const int N = ps.N[side], M = ps.M[side];
    const double* __restrict__ S = ps.S[side];
    const double* __restrict__ V = ps.V[side];
    double* __restrict__ Out = ps.Out[side];
    const int KS = tk_kept_entries(topk, VPT);
    KeptEntry* kept = reinterpret_cast<KeptEntry*>(tk_smem) + (size_t)warp * KS;
    KeptEntry* alist = reinterpret_cast<KeptEntry*>(tk_smem) + (size_t)WARPS * KS + (size_t)warp * TK_LMAX;
    double* etab = reinterpret_cast<double*>(reinterpret_cast<KeptEntry*>(tk_smem) + (size_t)WARPS * (KS + TK_LMAX));
    double* sV = etab + 64;                                               // SMEM_V: [M][LDH_V]
    __shared__ __align__(8) uint64_t v_bar;
    exp_table_to_shared(etab);
    long long bh; int i, iend;
    if (SMEM_V) {
        bh = blockIdx.x;
        const int per = (N + (int)gridDim.y - 1) / (int)gridDim.y;
        i = (int)blockIdx.y * per + warp;
        iend = min(N, ((int)blockIdx.y + 1) * per);
        if (threadIdx.x == 0) { mbar_init(&v_bar, 1); mbar_fence_init(); }
        __syncthreads();
        pdl_wait();
        pdl_trigger();
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)((size_t)M * LDH_V * sizeof(double));
            mbar_expect_tx(&v_bar, bytes);
            bulk_g2s(sV, V + bh * (long long)M * LDH_V, bytes, &v_bar);
        }
    } else {
        __syncthreads();
        pdl_wait();
        pdl_trigger();
        const long long g0 = (long long)blockIdx.x * TK_WARPS + warp;      // row in (B,4,N) order
        bh = g0 / N;
        i = (int)(g0 - bh * N);
        iend = N;
        if (bh >= nbh) return;                                             // past the last row of this side (whole warps)
    }
    const int b = (int)(bh / HEADS), h = (int)(bh - (long long)b * HEADS);
    auto load_row = [&](int row, double (&dst)[VPT]) {
        const double* srow = S + (bh * N + row) * (long long)M;
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const int j = lane + 32 * v;
            dst[v] = (FULLROW || j < M) ? srow[j] : -INFINITY;          // padding: bin 0, never kept
        }
    };
    double s[VPT];
    if (i < iend) load_row(i, s);
    bool v_ready = !SMEM_V;
    // P.V lane roles: quarter-warp qw takes the kept entries t = qw (mod 4); lane cl of it the channels 2cl, 2cl+1,
    // 16+2cl, 17+2cl -- its two 16-byte reads of a value row fall on two contiguous 128-byte runs per quarter-warp
    const int qw = lane >> 3, cl = lane & 7;
    const uint32_t kq_s = (uint32_t)__cvta_generic_to_shared(kept) + (uint32_t)qw * 16u;         // entry qw + 4 n at kq_s + 64 n
    const uint32_t vq_s = (uint32_t)__cvta_generic_to_shared(sV) + (uint32_t)cl * 16u;
    const char* vq_g = reinterpret_cast<const char*>(V + bh * (long long)M * LDH_V) + cl * 16;
    (void)qw;
  for (; i < iend; i += WARPS) {
    double mx;
    if (!topk_select_fast<VPT, FULLROW>(s, M, topk, lane, kept, alist, mx))
        topk_select_exact<VPT>(S + (bh * N + i) * (long long)M, M, topk, lane, kept);
    const int nk = min(topk, M);
    // the logits of this warp's next row travel while the P.V below runs
    if (SMEM_V && i + WARPS < iend) load_row(i + WARPS, s);
    const double sum = topk_softmax_kept(kept, nk, mx, etab, lane);
    if (!v_ready) { mbar_wait(&v_bar, 0); v_ready = true; }
    topk_pv_store<SMEM_V>(kq_s, vq_s, vq_g, nk, sum, lane, Out + ((long long)b * N + i) * ldo + h * HDIM);
    __syncwarp();                                        // kept[] is rewritten by the next row
    if (!SMEM_V) break;
  }
}

// ------------------------------------------------------------------------------------------
// dynamic_attention() in ONE kernel: the dense logits never travel to HBM.
//
// The two-kernel route above writes the (B,4,N,M) float64 logits (537 MB per layer at cfg2) and reads them back: the
// logits kernel is bound by the HBM WRITE rate (2.8 TB/s, not by its DMMAs), the selection kernel by the shared-memory
// wavefronts of its value gathers -- the FP64 pipe idles there. Here both run in the same persistent CTA (one per SM):
//   warps 0-15   consumers: exact top-k selection + softmax + sparse P.V, one query row per warp (code shared with
//                topk_softmax_pv_kernel)
//   warps 16-19  producers: Q K^T / sqrt(32) of 16-row groups on DMMA.8x8x4 (one producer warp per SM sub-partition),
//                K arriving in 64-key chunks by TMA bulk copies; a finished group (16 rows x M logits) goes to a
//                slot of the CTA's private ring in global memory
//   warp 20      loader: the K chunk ring and the head's value matrix (one bulk copy per (side, b, h))
// The ring (8 slots x 16 rows x M doubles = 512 KB per CTA, 76 MB for 148 CTAs) is rewritten every few microseconds
// and stays in the 126 MB L2: the consumers read it with ld.global.cg. A CTA walks a contiguous range of the work
// items (side, b, h, row split), so the value matrix is reloaded only when (side, b, h) changes.
// ------------------------------------------------------------------------------------------
constexpr int TF_CONS = 16, TF_PROD = 4, TF_THREADS = 32 * (TF_CONS + TF_PROD), TF_GROUP = 16, TF_SLOTS = 8, TF_KC = 64;
struct TopkFusedParams {
    const double* Q[2]; const double* K[2]; const double* V[2]; double* Out[2];
    int N[2], M[2];
    double* ring;                 // [gridDim.x][TF_SLOTS][TF_GROUP][ldS]
    int ldS, ldo, topk, nbh, nsplit, per, items, items_side0, slots;   // slots <= TF_SLOTS ring slots in use
    double scale;
};

template <bool FULLROW>
__global__ void __launch_bounds__(TF_THREADS, 1) topk_fused_kernel(const __grid_constant__ TopkFusedParams p) {
    constexpr int VPT = 16;
    extern __shared__ __align__(16) unsigned char tf_smem_raw[];
    constexpr uint32_t HB = 64 * VPT;
    const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(tf_smem_raw);
    unsigned char* smem = tf_smem_raw + (((raw_s + HB - 1) & ~(HB - 1)) - raw_s);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int KS = tk_kept_entries(p.topk, VPT);
    KeptEntry* kept_all = reinterpret_cast<KeptEntry*>(smem);
    KeptEntry* alist_all = kept_all + (size_t)TF_CONS * KS;
    double* etab = reinterpret_cast<double*>(alist_all + (size_t)TF_CONS * TK_LMAX);
    double* sK = etab + 64;                                            // [2][TF_KC][LDH_QK]
    double* sV = sK + 2 * TF_KC * LDH_QK;                              // [M][LDH_V]
    __shared__ __align__(8) uint64_t k_full[2], k_empty[2], v_full, slot_full[TF_SLOTS], slot_free[TF_SLOTS];
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], TF_PROD); }
        mbar_init(&v_full, 1);
        for (int i = 0; i < TF_SLOTS; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_free[i], TF_CONS); }
        mbar_fence_init();
    }
    exp_table_to_shared(etab);
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    // contiguous item range of this CTA; item = ((side * nbh + bh) * nsplit + split)
    const int it0 = (int)(((long long)p.items * blockIdx.x) / gridDim.x), it1 = (int)(((long long)p.items * (blockIdx.x + 1)) / gridDim.x);
    double* ring = p.ring + (size_t)blockIdx.x * p.slots * TF_GROUP * p.ldS;
    auto decode = [&](int it, int& side, int& bh, int& r0, int& r1) {
        side = it >= p.items_side0 ? 1 : 0;
        const int q = it - side * p.items_side0;
        bh = q / p.nsplit;
        const int sp = q - bh * p.nsplit;
        r0 = sp * p.per;
        r1 = min(p.N[side], r0 + p.per);
    };

    if (warp >= TF_CONS) {
        // ------------------------------------------------------------------ producers: logits of 16-row groups
        const int pw = warp - TF_CONS, qr = lane >> 2, qc = lane & 3;
        const bool loader = pw == 0 && elect_one();          // one lane of producer warp 0 also feeds the K chunk ring
        // the loader walks the chunk sequence (item, pass, chunk) one chunk ahead of the compute loop below
        int l_it = it0, l_ps = 0, l_c = 0, l_q = 0;
        auto load_next = [&]() {                              // issues the load of chunk l_q (if any is left)
            int side, bh, r0, r1;
            for (;; ++l_it) {                                 // items past the end of the shorter side are empty
                if (l_it >= it1) return;
                decode(l_it, side, bh, r0, r1);
                if (r1 > r0) break;
            }
            const int M = p.M[side];
            const int groups = (r1 - r0 + TF_GROUP - 1) / TF_GROUP, passes = (groups + TF_PROD - 1) / TF_PROD;
            const int nch = (M + TF_KC - 1) / TF_KC;
            const int st = l_q & 1;
            if (l_q >= 2) mbar_wait(&k_empty[st], (unsigned)(((l_q >> 1) - 1) & 1));
            const int rows = min(TF_KC, M - l_c * TF_KC);
            const unsigned bytes = (unsigned)(rows * LDH_QK * sizeof(double));
            mbar_expect_tx(&k_full[st], bytes);
            bulk_g2s(sK + st * TF_KC * LDH_QK, p.K[side] + ((long long)bh * M + (long long)l_c * TF_KC) * LDH_QK, bytes, &k_full[st]);
            ++l_q;
            if (++l_c == nch) { l_c = 0; if (++l_ps == passes) { l_ps = 0; ++l_it; } }
        };
        if (loader) load_next();
        int kq = 0, uses[2] = {0, 0};                       // uses[j]: how often this warp has filled slot pw + 4 j
        for (int it = it0; it < it1; ++it) {
            int side, bh, r0, r1;
            decode(it, side, bh, r0, r1);
            if (r1 <= r0) continue;
            const int N = p.N[side], M = p.M[side];
            const int groups = (r1 - r0 + TF_GROUP - 1) / TF_GROUP, passes = (groups + TF_PROD - 1) / TF_PROD;
            const int nch = (M + TF_KC - 1) / TF_KC;
            const double* Qbh = p.Q[side] + (long long)bh * N * LDH_QK;
            for (int ps = 0; ps < passes; ++ps) {
                const int g = ps * TF_PROD + pw;            // group of this warp in this pass (slot g)
                const bool active = g < groups;
                double qa[2][8];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int row = r0 + g * TF_GROUP + mt * 8 + qr;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) qa[mt][ks] = (active && row < N) ? Qbh[(long long)row * LDH_QK + ks * 4 + qc] : 0.0;
                }
                if (active) {
                    const int use = uses[ps & 1];           // at most two passes per item (8 slots of 16 rows)
                    if (use > 0) mbar_wait(&slot_free[g], (unsigned)((use - 1) & 1));
                }
                double* out = ring + (size_t)(active ? g : 0) * TF_GROUP * p.ldS;
                for (int c = 0; c < nch; ++c, ++kq) {
                    const int st = kq & 1;
                    if (loader) load_next();                 // chunk kq + 1 travels while chunk kq is multiplied
                    mbar_wait(&k_full[st], (unsigned)((kq >> 1) & 1));
                    if (active) {
                        const double* ks_ = sK + st * TF_KC * LDH_QK;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            double acc[2][4][2];
#pragma unroll
                            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                                for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                                for (int nt = 0; nt < 4; ++nt) {
                                    const double bfrag = ks_[(half * 32 + nt * 8 + qr) * LDH_QK + ks * 4 + qc];
                                    dmma884(acc[0][nt][0], acc[0][nt][1], qa[0][ks], bfrag);
                                    dmma884(acc[1][nt][0], acc[1][nt][1], qa[1][ks], bfrag);
                                }
                            }
#pragma unroll
                            for (int mt = 0; mt < 2; ++mt) {
                                double* orow = out + (size_t)(mt * 8 + qr) * p.ldS;
#pragma unroll
                                for (int nt = 0; nt < 4; ++nt) {
                                    const int col = c * TF_KC + half * 32 + nt * 8 + 2 * qc;
                                    const double y0 = acc[mt][nt][0] * p.scale, y1 = acc[mt][nt][1] * p.scale;
                                    if (col + 1 < p.ldS) *reinterpret_cast<double2*>(orow + col) = make_double2(y0, y1);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&k_empty[st]);
                }
                if (active) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&slot_full[g]);      // release: the group's logits are visible to the consumers
                    ++uses[ps & 1];
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        KeptEntry* kept = kept_all + (size_t)warp * KS;
        KeptEntry* alist = alist_all + (size_t)warp * TK_LMAX;
        const int qw = lane >> 3, cl = lane & 7;
        const uint32_t kq_s = (uint32_t)__cvta_generic_to_shared(kept) + (uint32_t)qw * 16u;
        const uint32_t vq_s = (uint32_t)__cvta_generic_to_shared(sV) + (uint32_t)cl * 16u;
        int uses[TF_SLOTS];
#pragma unroll
        for (int i = 0; i < TF_SLOTS; ++i) uses[i] = 0;
        int prev_side = -1, prev_bh = -1;
        unsigned vphase = 0;
        for (int it = it0; it < it1; ++it) {
            int side, bh, r0, r1;
            decode(it, side, bh, r0, r1);
            if (r1 <= r0) continue;
            const int N = p.N[side], M = p.M[side];
            const int b = bh / HEADS, h = bh - b * HEADS;
            const int groups = (r1 - r0 + TF_GROUP - 1) / TF_GROUP;
            // the head's value matrix: reloaded when (side, b, h) changes, once every consumer is done with the previous item
            if (side != prev_side || bh != prev_bh) {
                asm volatile("bar.sync 1, %0;" :: "n"(32 * TF_CONS) : "memory");
                if (warp == 0 && elect_one()) {
                    const unsigned bytes = (unsigned)((size_t)M * LDH_V * sizeof(double));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic reads of the old matrix before the async write
                    mbar_expect_tx(&v_full, bytes);
                    bulk_g2s(sV, p.V[side] + (long long)bh * M * LDH_V, bytes, &v_full);
                }
                mbar_wait(&v_full, vphase);
                vphase ^= 1u;
                prev_side = side; prev_bh = bh;
            }
#pragma unroll
            for (int g = 0; g < TF_SLOTS; ++g) {
                if (g < groups) {
                    mbar_wait(&slot_full[g], (unsigned)(uses[g] & 1));
                    ++uses[g];
                    const int row = r0 + g * TF_GROUP + warp;
                    if (row < r1) {
                        const double* srow = ring + ((size_t)g * TF_GROUP + warp) * p.ldS;
                        double s[VPT];
#pragma unroll
                        for (int v = 0; v < VPT; ++v) {
                            const int j = lane + 32 * v;
                            s[v] = (FULLROW || j < M) ? __ldcg(srow + j) : -INFINITY;
                        }
                        double mx;
                        if (!topk_select_fast<VPT, FULLROW>(s, M, p.topk, lane, kept, alist, mx))
                            topk_select_exact<VPT>(srow, M, p.topk, lane, kept);
                        const int nk = min(p.topk, M);
                        const double sum = topk_softmax_kept(kept, nk, mx, etab, lane);
                        topk_pv_store<true>(kq_s, vq_s, nullptr, nk, sum, lane, p.Out[side] + ((long long)b * N + row) * p.ldo + h * HDIM);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&slot_free[g]);
                }
            }
        }
    }
}

static size_t topk_fused_smem(int topk, int mmax) {
    return ((size_t)TF_CONS * (tk_kept_entries(topk, 16) + TK_LMAX)) * sizeof(KeptEntry) + 64 * sizeof(double) +
           (size_t)2 * TF_KC * LDH_QK * sizeof(double) + (size_t)mmax * LDH_V * sizeof(double) + 64 * 16;
}

static int tf_sms() {
    static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n > 0 ? n : 148; }();
    return sms;
}
// items and ring geometry shared by the sizing function and the launcher
// ring slots in use (MDGAT_TOPK_SLOTS = 4 | 8): 8 double-buffers the four producers, 4 halves the ring (38 MB for 148 CTAs at M = 512)
static int tf_slots() {
    static const int n = [] { const char* e = getenv("MDGAT_TOPK_SLOTS"); return e && e[0] == '4' ? 4 : TF_SLOTS; }();
    return n;
}
static void tf_geometry(int nmax, int mmax, int& nsplit, int& per, int& ldS) {
    const int ns0 = (nmax + tf_slots() * TF_GROUP - 1) / (tf_slots() * TF_GROUP);
    per = ((nmax + ns0 - 1) / ns0 + TF_GROUP - 1) / TF_GROUP * TF_GROUP;
    nsplit = (nmax + per - 1) / per;
    ldS = (mmax + 1) & ~1;
}
// ring scratch (doubles) that covers every top-k layer of a forward on B pairs of N x M keypoints (self and cross layers)
size_t topk_fused_ring_doubles(int B, int N, int M) {
    const int big = N > M ? N : M;
    int nsplit, per, ldS;
    tf_geometry(big, big, nsplit, per, ldS);
    const long long items = 2ll * B * HEADS * nsplit;
    const long long grid = items < tf_sms() ? items : tf_sms();
    return (size_t)grid * tf_slots() * TF_GROUP * ldS;
}
// exact ring need of one launch
size_t topk_fused_ring_need(const AttnSides& ps, int B, int nsides) {
    int nmax = 0, mmax = 0;
    for (int s = 0; s < nsides; ++s) { nmax = ps.N[s] > nmax ? ps.N[s] : nmax; mmax = ps.M[s] > mmax ? ps.M[s] : mmax; }
    int nsplit, per, ldS;
    tf_geometry(nmax, mmax, nsplit, per, ldS);
    const long long items = (long long)nsides * B * HEADS * nsplit;
    const long long grid = items < tf_sms() ? items : tf_sms();
    return (size_t)grid * tf_slots() * TF_GROUP * ldS;
}
bool topk_fused_supported(int nsides, const int* N, const int* M, int topk) {
    int mmax = 0, nmin = 1 << 30;
    for (int s = 0; s < nsides; ++s) { mmax = M[s] > mmax ? M[s] : mmax; nmin = N[s] < nmin ? N[s] : nmin; }
    return mmax <= 512 && nmin >= 1 && topk >= 1 && topk_fused_smem(topk, mmax) <= 227 * 1024;
}

cudaError_t launch_topk_fused(const AttnSides& ps, int B, int nsides, int ldo, int topk, double* ring, size_t ring_doubles, cudaStream_t st) {
    const int sms = tf_sms();
    TopkFusedParams p;
    int mmax = 0, nmax = 0;
    bool full = true;
    for (int s = 0; s < 2; ++s) {
        const int t = s < nsides ? s : 0;
        p.Q[s] = ps.Q[t]; p.K[s] = ps.K[t]; p.V[s] = ps.V[t]; p.Out[s] = ps.Out[t]; p.N[s] = ps.N[t]; p.M[s] = ps.M[t];
        mmax = ps.M[t] > mmax ? ps.M[t] : mmax; nmax = ps.N[t] > nmax ? ps.N[t] : nmax;
        full = full && ps.M[t] == 512;
    }
    if (!topk_fused_supported(nsides, ps.N, ps.M, topk)) return cudaErrorInvalidValue;
    p.ring = ring; p.ldo = ldo; p.topk = topk; p.nbh = B * HEADS;
    // rows of a (side, b, h) per item: at most 8 groups of 16 (the ring), a multiple of 16
    tf_geometry(nmax, mmax, p.nsplit, p.per, p.ldS);
    p.items_side0 = p.nbh * p.nsplit;
    p.items = nsides * p.items_side0;
    p.scale = 1.0 / sqrt((double)HDIM);
    const size_t smem = topk_fused_smem(topk, mmax);
    const int grid = p.items < sms ? p.items : sms;
    p.slots = tf_slots();
    if ((size_t)grid * p.slots * TF_GROUP * p.ldS > ring_doubles) return cudaErrorInvalidValue;
    cudaError_t e;
    if (full) {
        if ((e = cudaFuncSetAttribute(topk_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        e = launch_pdl(topk_fused_kernel<true>, dim3(grid), dim3(TF_THREADS), smem, st, p);
    } else {
        if ((e = cudaFuncSetAttribute(topk_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        e = launch_pdl(topk_fused_kernel<false>, dim3(grid), dim3(TF_THREADS), smem, st, p);
    }
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Exact top-k THRESHOLD of every logits row, for the tcgen05 top-k attention (attention_i8.cu, TOPK mode): the masked
// softmax . V of dynamic_attention() (mdgat.py:196-210) then runs on the tensor cores as a full attention whose
// probabilities outside the kept set are zero. One warp per (b, h, query) row, two rows per warp in flight.
// Per row: thr and jlast such that the kept set is { j : z_j > thr  or  (z_j == thr and j <= jlast) } -- exactly k
// entries, ties at the k-th value resolved towards the lowest index (torch.topk keeps exactly k; which of several equal
// logits it keeps is unspecified, duplicated keypoints give identical value rows so the message does not depend on it)
// -- and the exact row maximum (the softmax shift: the kept set always contains it).
// The comparison is on the float64 logits bit for bit as the LOGITS mode of attn_i8_kernel stored them; the TOPK mode
// recomputes them with the same instruction sequence, so the two kernels agree on every entry.
// ------------------------------------------------------------------------------------------
template <int VPT>
__global__ void __launch_bounds__(32 * TK_WARPS)
topk_threshold_kernel(const double* __restrict__ S, double* __restrict__ thr_out, int* __restrict__ jlast_out,
                      double* __restrict__ max_out, int M, int topk, long long total_rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * TK_WARPS + warp;
    if (row >= total_rows) return;
    const double* srow = S + row * (long long)M;
    double s[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const int j = lane + 32 * v;
        s[v] = j < M ? srow[j] : -INFINITY;                 // padding never passes a ">= finite" test
    }
    double mx = -INFINITY, mn = INFINITY;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        if (lane + 32 * v < M) { mx = fmax(mx, s[v]); mn = fmin(mn, s[v]); }
    }
    mx = warp_max_d(mx);
    mn = -warp_max_d(-mn);
    double thr = mn;
    int jlast = 0x7fffffff;
    bool exact = topk >= M;
    if (!exact) {
        // bracket [lo, hi] with count(>= lo) = clo > k > chi = count(>= hi); even steps place the probe by linear
        // interpolation of the counts, odd steps bisect (see topk_softmax_pv_kernel)
        double lo = mn, hi = mx;
        int clo = M, chi = 1;
        for (int step = 0; step < 96; ++step) {
            double mid = lo + 0.5 * (hi - lo);
            if (!(mid > lo && mid < hi)) break;             // interval collapsed: ties or adjacent doubles
            if ((step & 1) == 0) {
                const double guess = lo + (hi - lo) * ((double)(clo - topk) / (double)(clo - chi));
                if (guess > lo && guess < hi) mid = guess;
            }
            int c = 0;
#pragma unroll
            for (int v = 0; v < VPT; ++v) c += (s[v] >= mid) ? 1 : 0;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c == topk) { thr = mid; exact = true; break; }
            if (c > topk) { lo = mid; clo = c; } else { hi = mid; chi = c; }
        }
    }
    if (!exact) {
        // ties at the k-th value (or adjacent doubles): most-significant-bit-first search on the order-preserving
        // integer image, then the index of the last tied entry to keep
        auto keyof = [&](int v) -> unsigned long long { return (lane + 32 * v) < M ? order_key(s[v]) : 0ull; };
        const unsigned long long kmax = order_key(mx), kmin = order_key(mn);
        const unsigned long long diff = kmax ^ kmin;
        unsigned long long prefix = kmax;
        bool found = false;
        if (diff != 0ull) {
            const int top = 63 - __clzll((long long)diff);
            prefix = (top == 63) ? 0ull : (kmax >> (top + 1)) << (top + 1);
            for (int bit = top; bit >= 0; --bit) {
                const unsigned long long cand = prefix | (1ull << bit);
                int c = 0;
#pragma unroll
                for (int v = 0; v < VPT; ++v) c += (keyof(v) >= cand) ? 1 : 0;
                c = __reduce_add_sync(0xffffffffu, c);
                if (c >= topk) {
                    prefix = cand;
                    if (c == topk) { found = true; break; }
                }
            }
        }
        // the double whose key is `prefix`: z >= thr  <=>  key(z) >= prefix
        thr = __longlong_as_double((long long)((prefix >> 63) ? (prefix ^ 0x8000000000000000ull) : ~prefix));
        if (!found) {
            int gt = 0;
#pragma unroll
            for (int v = 0; v < VPT; ++v) gt += (keyof(v) > prefix) ? 1 : 0;
            gt = __reduce_add_sync(0xffffffffu, gt);
            const int need = topk - gt;                     // tied entries to keep, lowest index first
            int seen = 0, jl = -1;
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
                const bool eq = keyof(v) == prefix;
                const unsigned em = __ballot_sync(0xffffffffu, eq);
                if (eq && seen + __popc(em & ((1u << lane) - 1u)) == need - 1) jl = lane + 32 * v;
                seen += __popc(em);
            }
            jlast = __reduce_max_sync(0xffffffffu, jl);
        }
    }
    if (lane == 0) { thr_out[row] = thr; jlast_out[row] = jlast; max_out[row] = mx; }
}

cudaError_t launch_topk_threshold(const double* S, double* thr, int* jlast, double* rmax, int B, int N, int M, int topk,
                                  cudaStream_t st) {
    if (B <= 0 || N <= 0) return cudaSuccess;
    const long long rows = (long long)B * HEADS * N;
    const unsigned grid = (unsigned)((rows + TK_WARPS - 1) / TK_WARPS);
    if (M <= 512) topk_threshold_kernel<16><<<grid, 32 * TK_WARPS, 0, st>>>(S, thr, jlast, rmax, M, topk, rows);
    else if (M <= 1024) topk_threshold_kernel<32><<<grid, 32 * TK_WARPS, 0, st>>>(S, thr, jlast, rmax, M, topk, rows);
    else if (M <= 2048) topk_threshold_kernel<64><<<grid, 32 * TK_WARPS, 0, st>>>(S, thr, jlast, rmax, M, topk, rows);
    else return cudaErrorInvalidValue;
    count_launch();
    return cudaGetLastError();
}

static size_t topk_smem_bytes(int warps, int topk, int vpt, int m_smem_v) {
    return ((size_t)warps * (tk_kept_entries(topk, vpt) + TK_LMAX)) * sizeof(KeptEntry) + 64 * sizeof(double) +
           (size_t)m_smem_v * LDH_V * sizeof(double) + 64 * vpt;            // + alignment slack of the kept lists
}

template <int VPT, bool FULLROW>
static cudaError_t launch_topk_tf(const TopkSides& ps, int nsides, int ldo, int B, int topk, int nmax, int nmin, int mmax, cudaStream_t st) {
    const int nbh = B * HEADS;
    // value matrix of a head resident in shared memory when it fits next to the kept lists (M = 512, k = 128: 197 KB)
    const size_t smem_v = topk_smem_bytes(TKS_WARPS, topk, VPT, mmax);
    if constexpr (VPT == 16) {
        static const bool smem_variant = [] { const char* e = getenv("MDGAT_TOPK_SMEMV"); return !(e && e[0] == '0'); }();
        if (smem_variant && smem_v <= 227 * 1024 && nmin >= TKS_WARPS) {
            cudaError_t e = cudaFuncSetAttribute(topk_softmax_pv_kernel<VPT, true, FULLROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v);
            if (e != cudaSuccess) return e;
            // rows of a (side, b, h) over nsplit CTAs: the split that needs the fewest (fractional) waves of 148 CTAs; every
            // extra CTA reloads the value matrix, so ties go to the smaller split
            static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n > 0 ? n : 148; }();
            static const int split_env = [] { const char* e = getenv("MDGAT_TOPK_SPLIT"); return e ? atoi(e) : 0; }();
            int nsplit = 1; double best = 1e30;
            for (int ns = 1; ns <= 8; ns *= 2) {
                if (nmin / ns < TKS_WARPS) break;
                const long long ctas = (long long)nsides * nbh * ns;
                const double cost = (double)((ctas + sms - 1) / sms) / ns + 0.02 * ns;
                if (cost < best - 1e-9) { best = cost; nsplit = ns; }
            }
            if (split_env > 0 && nmin / split_env >= 1) nsplit = split_env;
            return launch_pdl(topk_softmax_pv_kernel<VPT, true, FULLROW>, dim3((unsigned)nbh, (unsigned)nsplit, (unsigned)nsides), dim3(32 * TKS_WARPS), smem_v, st, ps, ldo, topk, nbh);
        }
    }
    const size_t smem = topk_smem_bytes(TK_WARPS, topk, VPT, 0);
    cudaError_t e = cudaFuncSetAttribute(topk_softmax_pv_kernel<VPT, false, FULLROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)(((long long)nbh * nmax + TK_WARPS - 1) / TK_WARPS);
    return launch_pdl(topk_softmax_pv_kernel<VPT, false, FULLROW>, dim3(grid, 1, (unsigned)nsides), dim3(32 * TK_WARPS), smem, st, ps, ldo, topk, nbh);
}

template <int VPT>
static cudaError_t launch_topk_t(const TopkSides& ps, int nsides, int ldo, int B, int topk, cudaStream_t st) {
    int nmax = 0, nmin = 1 << 30, mmax = 0;
    bool full = true;                                        // every side has exactly 32 VPT sources: no padding tests
    for (int s = 0; s < nsides; ++s) {
        nmax = ps.N[s] > nmax ? ps.N[s] : nmax; nmin = ps.N[s] < nmin ? ps.N[s] : nmin; mmax = ps.M[s] > mmax ? ps.M[s] : mmax;
        full = full && ps.M[s] == 32 * VPT;
    }
    return full ? launch_topk_tf<VPT, true>(ps, nsides, ldo, B, topk, nmax, nmin, mmax, st)
                : launch_topk_tf<VPT, false>(ps, nsides, ldo, B, topk, nmax, nmin, mmax, st);
}

// both sides of a layer in one launch (S[s]: dense logits (B,4,N[s],M[s]); V[s]: head-major values of the source side)
cudaError_t launch_topk_softmax_pv_sides(const double* const* S, const double* const* V, double* const* Out, int ldo,
                                         int B, const int* N, const int* M, int nsides, int topk, cudaStream_t st) {
    if (B <= 0 || nsides <= 0) return cudaSuccess;
    TopkSides ps;
    int mmax = 0;
    for (int s = 0; s < 2; ++s) {
        const int t = s < nsides ? s : 0;
        ps.S[s] = S[t]; ps.V[s] = V[t]; ps.Out[s] = Out[t]; ps.N[s] = N[t]; ps.M[s] = M[t];
        if (N[t] <= 0 || M[t] <= 0) return cudaErrorInvalidValue;
        mmax = M[t] > mmax ? M[t] : mmax;
    }
    cudaError_t e;
    if (mmax <= 512) e = launch_topk_t<16>(ps, nsides, ldo, B, topk, st);
    else if (mmax <= 1024) e = launch_topk_t<32>(ps, nsides, ldo, B, topk, st);
    else if (mmax <= 2048) e = launch_topk_t<64>(ps, nsides, ldo, B, topk, st);
    else return cudaErrorInvalidValue;
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_topk_softmax_pv(const double* S, const double* V, double* Out, int ldo,
                                   int B, int N, int M, int topk, cudaStream_t st) {
    if (B <= 0 || N <= 0) return cudaSuccess;
    return launch_topk_softmax_pv_sides(&S, &V, &Out, ldo, B, &N, &M, 1, topk, st);
}

}  // namespace mdgat
