// Hand-written backward of attention() / dynamic_attention() (/root/reference/models/mdgat.py:190-210) for the training
// path. The reference keeps the (B,4,N,M) probabilities of every layer for autograd (2 x 268 MB per layer
// at batch 32, 2x512 keypoints); here only q, k, v and the message are kept and the probabilities are recomputed tile by
// tile (flash-attention style), in float64.
//
//   S = Q K^T / sqrt(32),  P = softmax over the kept set of S (all sources, or exactly the k largest per row),  O = P V
//   D_i = dO_i . O_i,  dP = dO V^T,  dS = P o (dP - D_i),  dQ = dS K / sqrt(32),  dK = dS^T Q / sqrt(32),  dV = P^T dO
//
// Two kernels over 64 x 64 tiles of one (b, h): attn_bwd_q_kernel owns 64 query rows (row statistics in a first sweep over
// the sources, dQ in a second), attn_bwd_kv_kernel owns 64 sources (dK, dV in one sweep over the queries). All tile
// products run on the FP64 tensor path (DMMA.8x8x4) from shared-memory tiles.
// Top-k layers: the kept set must be the forward's, ties included, so the logits are not recomputed tile-wise (a different
// summation order could move an entry across the threshold) but read from the dense logits the forward's own kernel
// writes (launch_attention_logits) together with the per-row threshold / last tied column of launch_topk_threshold.
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int AB_T = 64, AB_P = 33, AB_PT = 65;      // tile edge, pitch of the 32-wide operand tiles, pitch of the 64 x 64 tile
constexpr size_t AB_SMEM = ((size_t)4 * AB_T * AB_P + (size_t)AB_T * AB_PT + 3 * AB_T) * sizeof(double);

struct AttnBwdParams {
    const double* Q; const double* K; const double* V;     // head-major (B,4,n,ldq / ldq / ldv)
    const double* O; const double* dO;                     // (B,4,N,32)
    double* dQ; double* dK; double* dV;                    // (B,4,N,32), (B,4,M,32), (B,4,M,32)
    double* lse; double* dvec;                             // (B,4,N): log-sum-exp over the kept set, D_i
    const double* S; const double* thr; const int* jlast;  // top-k layers: dense logits (B,4,N,M), kept-set description; else null
    int N, M, ldq, ldv;
    double scale;
};

// rows [r0, r0 + 64) of a (n x 32) matrix with row pitch ld into a 64 x 33 tile, zero beyond n
DEVINL void ab_load_tile(double* dst, const double* src, int ld, int r0, int n, int tid) {
    for (int k = tid; k < AB_T * 32; k += 256) {
        const int r = k >> 5, c = k & 31;
        dst[r * AB_P + c] = (r0 + r < n) ? src[(long long)(r0 + r) * ld + c] : 0.0;
    }
}

// Tile products on the FP64 tensor path (mma.sync.m8n8k4.f64 -> DMMA.8x8x4). Warp w owns rows 8w .. 8w+7 of a tile; in a
// C fragment lane l holds row 8w + l/4 and the column pair 2 (l%4), 2 (l%4) + 1 of every 8-column block.
// c[nt][e] = (A B^T)[8w + l/4][8 nt + 2 (l%4) + e] for two 64 x 32 tiles A, B (pitch AB_P)
DEVINL void ab_dmma_nt(const double* A, const double* B, int warp, int lane, double (&c)[8][2]) {
    const int qr = lane >> 2, qc = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) c[nt][0] = c[nt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const double a = A[(8 * warp + qr) * AB_P + ks * 4 + qc];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) dmma884(c[nt][0], c[nt][1], a, B[(nt * 8 + qr) * AB_P + ks * 4 + qc]);
    }
}
// c[nt][e] += sum_k T'[8w + l/4][k] R[k][8 nt + 2 (l%4) + e] over the 64 k of a tile, R a 64 x 32 tile (pitch AB_P) and T' the
// 64 x 64 tile Tt (pitch AB_PT) read as stored (TRANS = false: T' = Tt) or transposed (TRANS = true: T' = Tt^T)
template <bool TRANS>
DEVINL void ab_dmma_acc(const double* Tt, const double* R, int warp, int lane, double (&c)[4][2]) {
    const int qr = lane >> 2, qc = lane & 3;
#pragma unroll 4
    for (int kk = 0; kk < 16; ++kk) {
        const int k = kk * 4 + qc;
        const double a = TRANS ? Tt[k * AB_PT + 8 * warp + qr] : Tt[(8 * warp + qr) * AB_PT + k];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(c[nt][0], c[nt][1], a, R[k * AB_P + nt * 8 + qr]);
    }
}

// scaled logits in C-fragment order: recomputed from the Q / K tiles, or read from the dense logits (top-k layers)
DEVINL void ab_logits(const AttnBwdParams& p, long long bh, const double* Qs, const double* Ks, int i0, int j0, int warp, int lane,
                      double (&s)[8][2]) {
    if (p.S == nullptr) {
        ab_dmma_nt(Qs, Ks, warp, lane, s);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { s[nt][0] *= p.scale; s[nt][1] *= p.scale; }
    } else {
        const int i = i0 + 8 * warp + (lane >> 2);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = j0 + nt * 8 + 2 * (lane & 3) + e;
                s[nt][e] = (i < p.N && j < p.M) ? p.S[(bh * p.N + i) * (long long)p.M + j] : 0.0;
            }
    }
}
// is source j of a query row in the softmax?
DEVINL bool ab_kept(const AttnBwdParams& p, double s, int j, double thr, int jl) {
    if (j >= p.M) return false;
    if (p.S == nullptr) return true;
    return s > thr || (s == thr && j <= jl);
}

// grid (ceil(N / 64), 4, B): row statistics, D_i and dQ of 64 query rows
__global__ void __launch_bounds__(256) attn_bwd_q_kernel(const AttnBwdParams p) {
    extern __shared__ __align__(16) double ab_sm[];
    double* Qs = ab_sm; double* Ks = Qs + AB_T * AB_P; double* Vs = Ks + AB_T * AB_P; double* dOs = Vs + AB_T * AB_P;
    double* Tt = dOs + AB_T * AB_P;                                  // [64][65] dS tile
    double* Dsm = Tt + AB_T * AB_PT; double* Lsm = Dsm + AB_T; double* Thr = Lsm + AB_T;
    __shared__ int Jl[AB_T];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, qr = lane >> 2, qc = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * AB_T;
    const long long bh = (long long)b * HEADS + h;
    const double* Qg = p.Q + bh * p.N * p.ldq;
    const double* Kg = p.K + bh * p.M * p.ldq;
    const double* Vg = p.V + bh * p.M * p.ldv;
    const double* Og = p.O + bh * p.N * 32;
    const double* dOg = p.dO + bh * p.N * 32;
    ab_load_tile(Qs, Qg, p.ldq, i0, p.N, tid);
    ab_load_tile(dOs, dOg, 32, i0, p.N, tid);
    if (tid < AB_T) {
        const int i = i0 + tid;
        double d = 0.0;
        if (i < p.N)
            for (int c = 0; c < 32; ++c) d = fma(dOg[(long long)i * 32 + c], Og[(long long)i * 32 + c], d);
        Dsm[tid] = d;
        Thr[tid] = (p.S != nullptr && i < p.N) ? p.thr[bh * p.N + i] : -INFINITY;
        Jl[tid] = (p.S != nullptr && i < p.N) ? p.jlast[bh * p.N + i] : 0x7fffffff;
        if (i < p.N) p.dvec[bh * p.N + i] = d;
    }
    __syncthreads();
    const int r = 8 * warp + qr;                                     // this thread's row of the tile (shared by its quad)
    const double thr = Thr[r]; const int jl = Jl[r];
    // ---- sweep 1: log-sum-exp over the kept set (online maximum; the four lanes of a quad hold the same m, l)
    double m = -INFINITY, l = 0.0;
    for (int j0 = 0; j0 < p.M; j0 += AB_T) {
        if (p.S == nullptr) { __syncthreads(); ab_load_tile(Ks, Kg, p.ldq, j0, p.M, tid); __syncthreads(); }
        double s[8][2];
        ab_logits(p, bh, Qs, Ks, i0, j0, warp, lane, s);
        double mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) if (ab_kept(p, s[nt][e], j0 + nt * 8 + 2 * qc + e, thr, jl)) mx = fmax(mx, s[nt][e]);
        mx = fmax(mx, shfl_xor_d(mx, 1));
        mx = fmax(mx, shfl_xor_d(mx, 2));
        const double mn = fmax(m, mx);
        double sum = 0.0;
        if (mn > -INFINITY) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) if (ab_kept(p, s[nt][e], j0 + nt * 8 + 2 * qc + e, thr, jl)) sum += exp(s[nt][e] - mn);
        }
        sum += shfl_xor_d(sum, 1);
        sum += shfl_xor_d(sum, 2);
        l = (m > -INFINITY ? l * exp(m - mn) : 0.0) + sum;
        m = mn;
    }
    const double lse = m + log(l);
    if (qc == 0 && i0 + r < p.N) p.lse[bh * p.N + i0 + r] = lse;
    const double dd = Dsm[r];
    // ---- sweep 2: dQ = sum_j dS_ij K_j
    double dq[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = 0.0;
    for (int j0 = 0; j0 < p.M; j0 += AB_T) {
        __syncthreads();
        ab_load_tile(Ks, Kg, p.ldq, j0, p.M, tid);
        ab_load_tile(Vs, Vg, p.ldv, j0, p.M, tid);
        __syncthreads();
        double s[8][2], dp[8][2];
        ab_logits(p, bh, Qs, Ks, i0, j0, warp, lane, s);
        ab_dmma_nt(dOs, Vs, warp, lane, dp);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = nt * 8 + 2 * qc + e;
                const double pr = ab_kept(p, s[nt][e], j0 + c, thr, jl) ? exp(s[nt][e] - lse) : 0.0;
                Tt[r * AB_PT + c] = pr * (dp[nt][e] - dd);
            }
        __syncthreads();
        ab_dmma_acc<false>(Tt, Ks, warp, lane, dq);
    }
    if (i0 + r < p.N) {
        double* dst = p.dQ + (bh * p.N + i0 + r) * 32 + 2 * qc;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { dst[nt * 8] = dq[nt][0] * p.scale; dst[nt * 8 + 1] = dq[nt][1] * p.scale; }
    }
}

// grid (ceil(M / 64), 4, B): dK and dV of 64 sources (needs lse and dvec from attn_bwd_q_kernel)
__global__ void __launch_bounds__(256) attn_bwd_kv_kernel(const AttnBwdParams p) {
    extern __shared__ __align__(16) double ab_sm[];
    double* Qs = ab_sm; double* Ks = Qs + AB_T * AB_P; double* Vs = Ks + AB_T * AB_P; double* dOs = Vs + AB_T * AB_P;
    double* Tt = dOs + AB_T * AB_P;
    double* Dsm = Tt + AB_T * AB_PT; double* Lsm = Dsm + AB_T; double* Thr = Lsm + AB_T;
    __shared__ int Jl[AB_T];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, qr = lane >> 2, qc = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * AB_T;
    const long long bh = (long long)b * HEADS + h;
    const double* Qg = p.Q + bh * p.N * p.ldq;
    const double* Kg = p.K + bh * p.M * p.ldq;
    const double* Vg = p.V + bh * p.M * p.ldv;
    const double* dOg = p.dO + bh * p.N * 32;
    ab_load_tile(Ks, Kg, p.ldq, j0, p.M, tid);
    ab_load_tile(Vs, Vg, p.ldv, j0, p.M, tid);
    // this thread's part of dK / dV: source 8 warp + qr, channels 8 nt + 2 qc, + 1
    double dk[4][2], dv[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) dk[nt][0] = dk[nt][1] = dv[nt][0] = dv[nt][1] = 0.0;
    const int r = 8 * warp + qr;                                     // query row of the logits tile this thread holds
    for (int i0 = 0; i0 < p.N; i0 += AB_T) {
        __syncthreads();
        ab_load_tile(Qs, Qg, p.ldq, i0, p.N, tid);
        ab_load_tile(dOs, dOg, 32, i0, p.N, tid);
        if (tid < AB_T) {
            const int i = i0 + tid;
            const bool ok = i < p.N;
            Dsm[tid] = ok ? p.dvec[bh * p.N + i] : 0.0;
            Lsm[tid] = ok ? p.lse[bh * p.N + i] : 0.0;
            Thr[tid] = (p.S != nullptr && ok) ? p.thr[bh * p.N + i] : -INFINITY;
            Jl[tid] = (p.S != nullptr && ok) ? p.jlast[bh * p.N + i] : 0x7fffffff;
        }
        __syncthreads();
        double s[8][2], dp[8][2], pr[8][2];
        ab_logits(p, bh, Qs, Ks, i0, j0, warp, lane, s);              // rows = queries, columns = sources
        ab_dmma_nt(dOs, Vs, warp, lane, dp);
        const bool rok = i0 + r < p.N;
        const double thr = Thr[r], lse = Lsm[r], dd = Dsm[r]; const int jl = Jl[r];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = nt * 8 + 2 * qc + e;
                pr[nt][e] = (rok && ab_kept(p, s[nt][e], j0 + c, thr, jl)) ? exp(s[nt][e] - lse) : 0.0;
                Tt[r * AB_PT + c] = pr[nt][e];
            }
        __syncthreads();
        ab_dmma_acc<true>(Tt, dOs, warp, lane, dv);                   // dV_j += sum_i P_ij dO_i
        __syncthreads();
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) Tt[r * AB_PT + nt * 8 + 2 * qc + e] = pr[nt][e] * (dp[nt][e] - dd);
        __syncthreads();
        ab_dmma_acc<true>(Tt, Qs, warp, lane, dk);                    // dK_j += sum_i dS_ij Q_i
    }
    const int j = j0 + 8 * warp + qr;
    if (j < p.M) {
        double* dkp = p.dK + (bh * p.M + j) * 32 + 2 * qc;
        double* dvp = p.dV + (bh * p.M + j) * 32 + 2 * qc;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            dkp[nt * 8] = dk[nt][0] * p.scale; dkp[nt * 8 + 1] = dk[nt][1] * p.scale;
            dvp[nt * 8] = dv[nt][0]; dvp[nt * 8 + 1] = dv[nt][1];
        }
    }
}

// scratch: lse (B 4 N) | dvec (B 4 N) | top-k only: thr (B 4 N) | rmax (B 4 N) | jlast ints (B 4 N) | logits (B 4 N M)
size_t attention_bwd_scratch_doubles(int B, int N, int M, int topk) {
    const size_t rows = (size_t)B * HEADS * N;
    return 2 * rows + (topk > 0 ? 2 * rows + (rows + 1) / 2 + rows * (size_t)M : 0) + 2;
}

cudaError_t launch_attention_backward(const double* Qh, const double* Kh, const double* Vh, const double* O, const double* dO,
                                      double* dQ, double* dK, double* dV, int B, int N, int M, int topk, double* scratch,
                                      cudaStream_t st) {
    const size_t rows = (size_t)B * HEADS * N;
    AttnBwdParams p;
    p.Q = Qh; p.K = Kh; p.V = Vh; p.O = O; p.dO = dO; p.dQ = dQ; p.dK = dK; p.dV = dV;
    p.lse = scratch; p.dvec = scratch + rows;
    p.S = nullptr; p.thr = nullptr; p.jlast = nullptr;
    p.N = N; p.M = M; p.ldq = LDH_QK; p.ldv = LDH_V; p.scale = 1.0 / sqrt((double)HDIM);
    cudaError_t e;
    if (topk > 0) {
        double* thr = scratch + 2 * rows;
        double* rmax = thr + rows;
        int* jlast = reinterpret_cast<int*>(rmax + rows);
        double* S = rmax + rows + (rows + 1) / 2;
        AttnSides ps;
        memset(&ps, 0, sizeof(ps));
        ps.Q[0] = Qh; ps.K[0] = Kh; ps.V[0] = Vh; ps.Out[0] = S; ps.N[0] = N; ps.M[0] = M;
        if ((e = launch_attention_logits(ps, B, 1, st)) != cudaSuccess) return e;
        if ((e = launch_topk_threshold(S, thr, jlast, rmax, B, N, M, topk, st)) != cudaSuccess) return e;
        p.S = S; p.thr = thr; p.jlast = jlast;
    }
    if ((e = cudaFuncSetAttribute(attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM)) != cudaSuccess) return e;
    attn_bwd_q_kernel<<<dim3((N + AB_T - 1) / AB_T, HEADS, B), 256, AB_SMEM, st>>>(p);
    attn_bwd_kv_kernel<<<dim3((M + AB_T - 1) / AB_T, HEADS, B), 256, AB_SMEM, st>>>(p);
    count_launch(2);
    return cudaGetLastError();
}

}  // namespace mdgat
