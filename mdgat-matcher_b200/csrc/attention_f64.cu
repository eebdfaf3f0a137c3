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
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int A_BM = 64, A_BN = 64, A_THREADS = 128, A_STAGES = 2;
constexpr size_t A_SMEM = (size_t)A_STAGES * A_BN * (LDH_QK + LDH_V) * sizeof(double);

__global__ void __launch_bounds__(A_THREADS, 2)
attn_full_kernel(const double* __restrict__ Q, const double* __restrict__ K, const double* __restrict__ V,
                 double* __restrict__ Out, int ldo, int N, int M, double scale) {
    extern __shared__ __align__(16) double smem[];
    double* Ks = smem;                                   // [stage][A_BN][LDH_QK]
    double* Vs = smem + A_STAGES * A_BN * LDH_QK;        // [stage][A_BN][LDH_V]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = blockIdx.y, b = blockIdx.z;
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

    // K/V chunks are contiguous in the head-major buffers (padded rows included): flat copy.
    auto load_chunk = [&](int c, int buf) {
        const int j0 = c * A_BN;
        const int rows = min(A_BN, M - j0);
        const double* ksrc = Kbh + (long long)j0 * LDH_QK;
        const double* vsrc = Vbh + (long long)j0 * LDH_V;
        double* kd = Ks + buf * A_BN * LDH_QK;
        double* vd = Vs + buf * A_BN * LDH_V;
        const int kvalid = rows * LDH_QK / 2, vvalid = rows * LDH_V / 2;     // 16-byte units
        for (int i = tid; i < A_BN * LDH_QK / 2; i += A_THREADS)
            cp_async16(kd + 2 * i, i < kvalid ? ksrc + 2 * i : ksrc, i < kvalid);
        for (int i = tid; i < A_BN * LDH_V / 2; i += A_THREADS)
            cp_async16(vd + 2 * i, i < vvalid ? vsrc + 2 * i : vsrc, i < vvalid);
    };

    load_chunk(0, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {
            load_chunk(c + 1, (c + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
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
        // columns past M (last chunk only)
        const int j0 = c * A_BN;
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
            const double alpha = exp((m_run[mt] - m_new) * scale);
            m_run[mt] = m_new;
            double rs = 0.0;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[mt][nt][0] = exp((s[mt][nt][0] - m_new) * scale);
                s[mt][nt][1] = exp((s[mt][nt][1] - m_new) * scale);
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

cudaError_t launch_attention_full(const double* Q, const double* K, const double* V, double* Out, int ldo,
                                  int B, int N, int M, cudaStream_t st) {
    if (B <= 0 || N <= 0 || M <= 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(attn_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A_SMEM);
    if (e != cudaSuccess) return e;
    dim3 grid((N + A_BM - 1) / A_BM, HEADS, B);
    attn_full_kernel<<<grid, A_THREADS, A_SMEM, st>>>(Q, K, V, Out, ldo, N, M, 1.0 / sqrt((double)HDIM));
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Exact top-k + softmax + sparse PV. One warp per (b, h, query) row; lane l holds logits
// j = l + 32 v. The k-th largest value is found by a most-significant-bit-first search on
// the order-preserving integer image of the doubles (at most 64 counting rounds, stops as
// soon as a candidate threshold keeps exactly k entries).
// ------------------------------------------------------------------------------------------
DEVINL unsigned long long order_key(double x) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    return (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
}

template <int VPT>
__global__ void __launch_bounds__(256)
topk_softmax_pv_kernel(const double* __restrict__ S, const double* __restrict__ V, double* __restrict__ Out,
                       int ldo, int N, int M, int topk, long long total_rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long g = (long long)blockIdx.x * 8 + warp;          // row in (B,4,N) order
    if (g >= total_rows) return;
    const long long bh = g / N;
    const int i = (int)(g - bh * N);
    const int b = (int)(bh / HEADS), h = (int)(bh - (long long)b * HEADS);
    const double* srow = S + g * (long long)M;
    const double* Vbh = V + bh * (long long)M * LDH_V;

    double s[VPT];
    double mx = -INFINITY;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const int j = lane + 32 * v;
        s[v] = j < M ? srow[j] : -INFINITY;
        mx = fmax(mx, s[v]);
    }
    mx = warp_max_d(mx);
    // entries past M get key 0 (below every real value, even -inf)
    auto keyof = [&](int v) -> unsigned long long { return (lane + 32 * v) < M ? order_key(s[v]) : 0ull; };

    unsigned long long prefix = 0ull;
    bool exact = false;
    for (int bit = 63; bit >= 0; --bit) {
        const unsigned long long cand = prefix | (1ull << bit);
        int c = 0;
#pragma unroll
        for (int v = 0; v < VPT; ++v) c += (keyof(v) >= cand) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= topk) {
            prefix = cand;
            if (c == topk) { exact = true; break; }
        }
    }
    // selection flags (bit v of sel)
    unsigned long long sel = 0ull;
    if (exact) {
#pragma unroll
        for (int v = 0; v < VPT; ++v) sel |= (unsigned long long)(keyof(v) >= prefix) << v;
    } else {
        // ties at the k-th value: keep everything above it plus the lowest-index equals
        int gt = 0;
#pragma unroll
        for (int v = 0; v < VPT; ++v) gt += (keyof(v) > prefix) ? 1 : 0;
        gt = __reduce_add_sync(0xffffffffu, gt);
        int need = topk - gt, seen = 0;
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const unsigned long long kv = keyof(v);
            const bool eq = (kv == prefix);
            const unsigned em = __ballot_sync(0xffffffffu, eq);
            const int rank = seen + __popc(em & ((1u << lane) - 1u));
            const bool take = (kv > prefix) || (eq && rank < need);
            sel |= (unsigned long long)take << v;
            seen += __popc(em);
        }
    }

    double sum = 0.0;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        s[v] = ((sel >> v) & 1ull) ? exp(s[v] - mx) : 0.0;
        sum += s[v];
    }
    sum = warp_sum_d(sum);

    double acc = 0.0;                                   // lane = channel d of this head
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        unsigned mask = __ballot_sync(0xffffffffu, (sel >> v) & 1ull);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const double pj = shfl_d(s[v], src);
            acc = fma(pj, __ldg(Vbh + (long long)(src + 32 * v) * LDH_V + lane), acc);
        }
    }
    Out[((long long)b * N + i) * ldo + h * HDIM + lane] = acc / sum;
}

cudaError_t launch_topk_softmax_pv(const double* S, const double* V, double* Out, int ldo,
                                   int B, int N, int M, int topk, cudaStream_t st) {
    if (B <= 0 || N <= 0) return cudaSuccess;
    const long long rows = (long long)B * HEADS * N;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (M <= 512) topk_softmax_pv_kernel<16><<<grid, 256, 0, st>>>(S, V, Out, ldo, N, M, topk, rows);
    else if (M <= 1024) topk_softmax_pv_kernel<32><<<grid, 256, 0, st>>>(S, V, Out, ldo, N, M, topk, rows);
    else if (M <= 2048) topk_softmax_pv_kernel<64><<<grid, 256, 0, st>>>(S, V, Out, ldo, N, M, topk, rows);
    else return cudaErrorInvalidValue;
    count_launch();
    return cudaGetLastError();
}

}  // namespace mdgat
