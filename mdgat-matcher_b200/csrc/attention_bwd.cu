// Hand-written backward of attention() / dynamic_attention() (/root/reference/models/mdgat.py:190-210) for the training
// path. The reference keeps the (B,4,N,M) probabilities of every layer for autograd (2 x 268 MB per layer
// at batch 32, 2x512 keypoints); here only q, k, v and the message are kept and the probabilities are recomputed tile by
// tile (flash-attention style), in float64.
//
//   S = Q K^T / sqrt(32),  P = softmax over the kept set of S (all sources, or exactly the k largest per row),  O = P V
//   D_i = dO_i . O_i,  dP = dO V^T,  dS = P o (dP - D_i),  dQ = dS K / sqrt(32),  dK = dS^T Q / sqrt(32),  dV = P^T dO
//
// Two kernels over 64 x 64 tiles of one (b, h): attn_bwd_q_kernel owns 64 query rows (row statistics in a first sweep over
// the sources, dQ in a second), attn_bwd_kv_kernel owns 64 sources (dK, dV in one sweep over the queries). Plain FP64 FMAs
// from shared-memory tiles: the training path is not the benchmarked one.
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

// this thread's 4 x 4 block (rows ty*4.., columns tx*4..) of A B^T for two 64 x 32 tiles
DEVINL void ab_mm_nt(const double* A, const double* B, int ty, int tx, double (&acc)[4][4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
    for (int d = 0; d < 32; ++d) {
        double x[4], y[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) x[a] = A[(ty * 4 + a) * AB_P + d];
#pragma unroll
        for (int b = 0; b < 4; ++b) y[b] = B[(tx * 4 + b) * AB_P + d];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(x[a], y[b], acc[a][b]);
    }
}

// scaled logits of this thread's 4 x 4 block: recomputed from the Q / K tiles, or read from the dense logits (top-k layers)
DEVINL void ab_logits(const AttnBwdParams& p, long long bh, const double* Qs, const double* Ks, int i0, int j0, int ty, int tx,
                      double (&s)[4][4]) {
    if (p.S == nullptr) {
        ab_mm_nt(Qs, Ks, ty, tx, s);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) s[a][b] *= p.scale;
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = i0 + ty * 4 + a;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = j0 + tx * 4 + b;
                s[a][b] = (i < p.N && j < p.M) ? p.S[(bh * p.N + i) * (long long)p.M + j] : 0.0;
            }
        }
    }
}
// is source j of query row i in the softmax?
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
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
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
    // ---- sweep 1: log-sum-exp over the kept set (online maximum; the 16 threads tx of a row group hold the same m, l)
    double m[4], l[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { m[a] = -INFINITY; l[a] = 0.0; }
    for (int j0 = 0; j0 < p.M; j0 += AB_T) {
        if (p.S == nullptr) { __syncthreads(); ab_load_tile(Ks, Kg, p.ldq, j0, p.M, tid); __syncthreads(); }
        double s[4][4];
        ab_logits(p, bh, Qs, Ks, i0, j0, ty, tx, s);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double thr = Thr[ty * 4 + a]; const int jl = Jl[ty * 4 + a];
            double mx = -INFINITY;
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) if (ab_kept(p, s[a][bb], j0 + tx * 4 + bb, thr, jl)) mx = fmax(mx, s[a][bb]);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmax(mx, shfl_xor_d(mx, o));
            const double mn = fmax(m[a], mx);
            double sum = 0.0;
            if (mn > -INFINITY) {
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) if (ab_kept(p, s[a][bb], j0 + tx * 4 + bb, thr, jl)) sum += exp(s[a][bb] - mn);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += shfl_xor_d(sum, o);
            l[a] = (m[a] > -INFINITY ? l[a] * exp(m[a] - mn) : 0.0) + sum;
            m[a] = mn;
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double lse = m[a] + log(l[a]);
            Lsm[ty * 4 + a] = lse;
            if (i0 + ty * 4 + a < p.N) p.lse[bh * p.N + i0 + ty * 4 + a] = lse;
        }
    }
    __syncthreads();
    // ---- sweep 2: dQ = sum_j dS_ij K_j
    double dq[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) dq[a][0] = dq[a][1] = 0.0;
    for (int j0 = 0; j0 < p.M; j0 += AB_T) {
        __syncthreads();
        ab_load_tile(Ks, Kg, p.ldq, j0, p.M, tid);
        ab_load_tile(Vs, Vg, p.ldv, j0, p.M, tid);
        __syncthreads();
        double s[4][4], dp[4][4];
        ab_logits(p, bh, Qs, Ks, i0, j0, ty, tx, s);
        ab_mm_nt(dOs, Vs, ty, tx, dp);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = ty * 4 + a;
            const double thr = Thr[r], lse = Lsm[r], dd = Dsm[r]; const int jl = Jl[r];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                const bool keep = ab_kept(p, s[a][bb], j0 + tx * 4 + bb, thr, jl);
                const double pr = keep ? exp(s[a][bb] - lse) : 0.0;
                Tt[r * AB_PT + tx * 4 + bb] = pr * (dp[a][bb] - dd);
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < AB_T; ++c) {
            const double k0 = Ks[c * AB_P + tx * 2], k1 = Ks[c * AB_P + tx * 2 + 1];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double ds = Tt[(ty * 4 + a) * AB_PT + c];
                dq[a][0] = fma(ds, k0, dq[a][0]);
                dq[a][1] = fma(ds, k1, dq[a][1]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        if (i < p.N) {
            double* dst = p.dQ + (bh * p.N + i) * 32 + tx * 2;
            dst[0] = dq[a][0] * p.scale;
            dst[1] = dq[a][1] * p.scale;
        }
    }
}

// grid (ceil(M / 64), 4, B): dK and dV of 64 sources (needs lse and dvec from attn_bwd_q_kernel)
__global__ void __launch_bounds__(256) attn_bwd_kv_kernel(const AttnBwdParams p) {
    extern __shared__ __align__(16) double ab_sm[];
    double* Qs = ab_sm; double* Ks = Qs + AB_T * AB_P; double* Vs = Ks + AB_T * AB_P; double* dOs = Vs + AB_T * AB_P;
    double* Tt = dOs + AB_T * AB_P;
    double* Dsm = Tt + AB_T * AB_PT; double* Lsm = Dsm + AB_T; double* Thr = Lsm + AB_T;
    __shared__ int Jl[AB_T];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * AB_T;
    const long long bh = (long long)b * HEADS + h;
    const double* Qg = p.Q + bh * p.N * p.ldq;
    const double* Kg = p.K + bh * p.M * p.ldq;
    const double* Vg = p.V + bh * p.M * p.ldv;
    const double* dOg = p.dO + bh * p.N * 32;
    ab_load_tile(Ks, Kg, p.ldq, j0, p.M, tid);
    ab_load_tile(Vs, Vg, p.ldv, j0, p.M, tid);
    // this thread's part of dK / dV: sources ty*4.., channels tx*2, tx*2+1
    double dk[4][2], dv[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) dk[a][0] = dk[a][1] = dv[a][0] = dv[a][1] = 0.0;
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
        double s[4][4], dp[4][4], pr[4][4];
        ab_logits(p, bh, Qs, Ks, i0, j0, ty, tx, s);           // rows = queries ty*4.., columns = sources tx*4..
        ab_mm_nt(dOs, Vs, ty, tx, dp);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = ty * 4 + a;
            const bool rok = i0 + r < p.N;
            const double thr = Thr[r], lse = Lsm[r]; const int jl = Jl[r];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                const bool keep = rok && ab_kept(p, s[a][bb], j0 + tx * 4 + bb, thr, jl);
                pr[a][bb] = keep ? exp(s[a][bb] - lse) : 0.0;
                Tt[r * AB_PT + tx * 4 + bb] = pr[a][bb];
            }
        }
        __syncthreads();
        // dV_j += sum_i P_ij dO_i
#pragma unroll 4
        for (int q = 0; q < AB_T; ++q) {
            const double g0 = dOs[q * AB_P + tx * 2], g1 = dOs[q * AB_P + tx * 2 + 1];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double pp = Tt[q * AB_PT + ty * 4 + a];
                dv[a][0] = fma(pp, g0, dv[a][0]);
                dv[a][1] = fma(pp, g1, dv[a][1]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = ty * 4 + a;
            const double dd = Dsm[r];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) Tt[r * AB_PT + tx * 4 + bb] = pr[a][bb] * (dp[a][bb] - dd);
        }
        __syncthreads();
        // dK_j += sum_i dS_ij Q_i
#pragma unroll 4
        for (int q = 0; q < AB_T; ++q) {
            const double q0 = Qs[q * AB_P + tx * 2], q1 = Qs[q * AB_P + tx * 2 + 1];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double ds = Tt[q * AB_PT + ty * 4 + a];
                dk[a][0] = fma(ds, q0, dk[a][0]);
                dk[a][1] = fma(ds, q1, dk[a][1]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int j = j0 + ty * 4 + a;
        if (j < p.M) {
            double* dkp = p.dK + (bh * p.M + j) * 32 + tx * 2;
            double* dvp = p.dV + (bh * p.M + j) * 32 + tx * 2;
            dkp[0] = dk[a][0] * p.scale; dkp[1] = dk[a][1] * p.scale;
            dvp[0] = dv[a][0]; dvp[1] = dv[a][1];
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
