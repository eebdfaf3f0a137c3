// Hand-written backward of log_optimal_transport (/root/reference/models/mdgat.py:279-308) for the training path
// (SURVEY.md 8 f-3). The reference back-propagates through 2 T unrolled logsumexp half-iterations with autograd, which
// retains 2 T tensors of shape (B, N+1, M+1) (13 GB at batch 32, 2x512 keypoints, T = 100). Here the gradient is formed
// from the scaling-form iterates in O(T (N + M)) extra memory per pair.
//
// Forward (scaling form, see sinkhorn.cu): c_i = max_j C_ij, E_ij = exp(C_ij - c_i), b_0 = 1,
//     a_t = mu / (E b_{t-1}),   b_t = nu / (E^T a_t),   t = 1..T       (a = exp(u + c), b = exp(v))
//     Z = C + u_T + v_T - norm.
// With G = dL/dZ the reverse sweep over the half-iterations is
//     gu = rowsum(G), gv = colsum(G)
//     for t = T..1:   q_t  = gv b_t / nu            gu -= a_t (E q_t)            (through v_t = log nu - LSE_i(C + u_t))
//                     p_t  = gu a_t / mu            gv  = -b_{t-1} (E^T p_t)     (through u_t = log mu - LSE_j(C + v_{t-1}))
//                     gu   = 0
// and every half-iteration adds a rank-one term to the gradient of the couplings:
//     dL/dC = G - E o sum_t (a_t q_t^T + p_t b_{t-1}^T).
// So the sweep needs only matrix-vector products with E (like the forward) plus one rank-2T contraction at the end; the
// iterates a_t, b_t are recomputed here (2 T matrix-vector products), nothing but C has to be kept from the forward.
// Plain global-memory kernels (one launch per half-iteration): the training path is not the benchmarked one.
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

// one warp per row: c_i, E row, range check (rows spanning more than SKB_MAX_RANGE cannot be carried in scaling form)
constexpr double SKB_MAX_RANGE = 600.0;
__global__ void __launch_bounds__(256)
skb_setup_kernel(const double* __restrict__ C, double* __restrict__ E, int* __restrict__ flag, int R1, int C1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, i = blockIdx.x * 8 + warp;
    if (i >= R1) return;
    const double* crow = C + ((long long)b * R1 + i) * C1;
    double* erow = E + ((long long)b * R1 + i) * C1;
    double mx = -INFINITY, mn = INFINITY;
    for (int j = lane; j < C1; j += 32) { const double c = crow[j]; mx = fmax(mx, c); mn = fmin(mn, c); }
    mx = warp_max_d(mx);
    mn = -warp_max_d(-mn);
    if (!(mx - mn < SKB_MAX_RANGE) && lane == 0) atomicOr(flag, 1);       // also NaN / inf
    for (int j = lane; j < C1; j += 32) erow[j] = exp(crow[j] - mx);
}

// Row products s_i = sum_j A_ij x_j (x == nullptr: x = 1), one warp per row, then per MODE:
//   RM_RAW    out_i = s_i
//   RM_FWD    out_i = mu_i / s_i                                              (a_t)
//   RM_BWD    gu_i = (gu_in ? gu_in_i : 0) - a_i s_i ;  out_i = gu_i a_i / mu_i   (p_t)
enum { RM_RAW = 0, RM_FWD = 1, RM_BWD = 2 };
template <int MODE>
__global__ void __launch_bounds__(256)
skb_rowmv_kernel(const double* __restrict__ A, const double* __restrict__ x, long long xstride, const double* __restrict__ a,
                 long long astride, const double* __restrict__ gu_in, double* __restrict__ out, long long ostride,
                 int R1, int C1, int N, int M) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, i = blockIdx.x * 8 + warp;
    if (i >= R1) return;
    const double* row = A + ((long long)b * R1 + i) * C1;
    const double* xb = x ? x + (long long)b * xstride : nullptr;
    double s = 0.0;
    for (int j = lane; j < C1; j += 32) s = fma(row[j], xb ? xb[j] : 1.0, s);
    s = warp_sum_d(s);
    if (lane != 0) return;
    const double mu = (i < N ? 1.0 : (double)M) / (double)(N + M);
    double r;
    if (MODE == RM_RAW) r = s;
    else if (MODE == RM_FWD) r = mu / s;
    else {
        const double ai = a[(long long)b * astride + i];
        const double gu = (gu_in ? gu_in[(long long)b * R1 + i] : 0.0) - ai * s;
        r = gu * ai / mu;
    }
    out[(long long)b * ostride + i] = r;
}

// Column products s_j = sum_i A_ij y_i (y == nullptr: y = 1), 32 columns x 16 row groups per CTA, fixed reduction order:
//   CM_RAW    out_j = s_j
//   CM_FWD    out_j = nu_j / s_j                                              (b_t)
//   CM_BWD    gv_j = -bprev_j s_j ;  out_j = gv_j bprev_j / nu_j                (q_{t-1})
//   CM_Q0     out_j = s_j bprev_j / nu_j with A = G, y = 1                      (q_T from gv_T = colsum(G), bprev = b_T)
enum { CM_RAW = 0, CM_FWD = 1, CM_BWD = 2, CM_Q0 = 3 };
constexpr int SKB_TY = 16;
template <int MODE>
__global__ void __launch_bounds__(32 * SKB_TY)
skb_colmv_kernel(const double* __restrict__ A, const double* __restrict__ y, long long ystride, const double* __restrict__ bprev,
                 long long bstride, double* __restrict__ out, long long ostride, int R1, int C1, int N, int M) {
    __shared__ double red[SKB_TY][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.y, j = blockIdx.x * 32 + tx;
    const double* Ab = A + (long long)b * R1 * C1;
    const double* yb = y ? y + (long long)b * ystride : nullptr;
    double s = 0.0;
    if (j < C1)
        for (int i = ty; i < R1; i += SKB_TY) s = fma(Ab[(long long)i * C1 + j], yb ? yb[i] : 1.0, s);
    red[ty][tx] = s;
    __syncthreads();
    if (ty != 0 || j >= C1) return;
#pragma unroll
    for (int k = 1; k < SKB_TY; ++k) s += red[k][tx];
    const double nu = (j < M ? 1.0 : (double)N) / (double)(N + M);
    double r;
    if (MODE == CM_RAW) r = s;
    else if (MODE == CM_FWD) r = nu / s;
    else {
        const double bp = bprev[(long long)b * bstride + j];
        r = (MODE == CM_BWD ? -bp * s : s) * bp / nu;
    }
    out[(long long)b * ostride + j] = r;
}

__global__ void skb_fill_kernel(double* __restrict__ p, double v, long long n, long long stride, int count) {
    // p[b * stride + k] = v for k < count
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long b = t / count, k = t - b * count;
    p[b * stride + k] = v;
}

// gC_ij = G_ij - E_ij sum_t (a_t,i q_t,j + p_t,i b_{t-1},j): 64 x 64 tile per CTA, 16 x 16 threads, 4 x 4 outputs each,
// the 2 T rank-one terms staged through shared memory 16 at a time.
//   ha: a_t at [b][t-1][R1] (t = 1..T)     hp: p_t at [b][t-1][R1]
//   hb: b_t at [b][t][C1]   (t = 0..T)     hq: q_t at [b][t-1][C1]
__global__ void __launch_bounds__(256)
skb_final_kernel(const double* __restrict__ G, const double* __restrict__ E, const double* __restrict__ ha,
                 const double* __restrict__ hp, const double* __restrict__ hb, const double* __restrict__ hq,
                 double* __restrict__ gC, int R1, int C1, int T) {
    __shared__ double sL[2][16][64], sR[2][16][64];       // [a | p][term][row], [q | b][term][col]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int b = blockIdx.z, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    for (int t0 = 0; t0 < T; t0 += 16) {
        __syncthreads();
        for (int k = threadIdx.x; k < 16 * 64; k += 256) {
            const int tt = k >> 6, l = k & 63, t = t0 + tt;
            const bool tv = t < T;
            const int i = i0 + l, j = j0 + l;
            sL[0][tt][l] = (tv && i < R1) ? ha[((long long)b * T + t) * R1 + i] : 0.0;
            sL[1][tt][l] = (tv && i < R1) ? hp[((long long)b * T + t) * R1 + i] : 0.0;
            sR[0][tt][l] = (tv && j < C1) ? hq[((long long)b * T + t) * C1 + j] : 0.0;
            sR[1][tt][l] = (tv && j < C1) ? hb[((long long)b * (T + 1) + t) * C1 + j] : 0.0;      // b_{t-1} of term t (1-based) = index t0-based
        }
        __syncthreads();
#pragma unroll 4
        for (int tt = 0; tt < 16; ++tt) {
            double la[4], lp[4], rq[4], rb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { la[r] = sL[0][tt][ty * 4 + r]; lp[r] = sL[1][tt][ty * 4 + r]; }
#pragma unroll
            for (int c = 0; c < 4; ++c) { rq[c] = sR[0][tt][tx * 4 + c]; rb[c] = sR[1][tt][tx * 4 + c]; }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(la[r], rq[c], fma(lp[r], rb[c], acc[r][c]));
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
        if (i >= R1) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx * 4 + c;
            if (j >= C1) continue;
            const long long o = ((long long)b * R1 + i) * C1 + j;
            gC[o] = G[o] - E[o] * acc[r][c];
        }
    }
}

// scratch: E (B R1 C1) | ha (B T R1) | hp (B T R1) | hb (B (T+1) C1) | hq (B T C1) | gu0 (B R1) | flag
size_t sinkhorn_bwd_scratch_doubles(int B, int N, int M, int iters) {
    const size_t R1 = N + 1, C1 = M + 1, T = iters > 0 ? iters : 1;
    return (size_t)B * (R1 * C1 + 2 * T * R1 + (T + 1) * C1 + T * C1 + R1) + 2;
}

cudaError_t launch_sinkhorn_backward(const double* C, const double* G, double* gC, double* scratch, int B, int N, int M,
                                     int iters, cudaStream_t st) {
    const int R1 = N + 1, C1 = M + 1, T = iters;
    const size_t n = (size_t)B * R1 * C1;
    if (T <= 0) return cudaMemcpyAsync(gC, G, n * sizeof(double), cudaMemcpyDeviceToDevice, st);     // Z = C - norm
    double* E = scratch;
    double* ha = E + n;
    double* hp = ha + (size_t)B * T * R1;
    double* hb = hp + (size_t)B * T * R1;
    double* hq = hb + (size_t)B * (T + 1) * C1;
    double* gu0 = hq + (size_t)B * T * C1;
    int* flag = reinterpret_cast<int*>(gu0 + (size_t)B * R1);
    cudaError_t e;
    if ((e = cudaMemsetAsync(flag, 0, sizeof(int), st)) != cudaSuccess) return e;
    const dim3 rgrid((R1 + 7) / 8, B), cgrid((C1 + 31) / 32, B), cblock(32, SKB_TY);
    const long long sa = (long long)T * R1, sb = (long long)(T + 1) * C1, sq = (long long)T * C1;
    skb_setup_kernel<<<rgrid, 256, 0, st>>>(C, E, flag, R1, C1);
    {   // b_0 = 1
        const long long tot = (long long)B * C1;
        skb_fill_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(hb, 1.0, tot, sb, C1);
    }
    // forward iterates: a_t at ha[t-1], b_t at hb[t]
    for (int t = 1; t <= T; ++t) {
        skb_rowmv_kernel<RM_FWD><<<rgrid, 256, 0, st>>>(E, hb + (size_t)(t - 1) * C1, sb, nullptr, 0, nullptr, ha + (size_t)(t - 1) * R1, sa, R1, C1, N, M);
        skb_colmv_kernel<CM_FWD><<<cgrid, cblock, 0, st>>>(E, ha + (size_t)(t - 1) * R1, sa, nullptr, 0, hb + (size_t)t * C1, sb, R1, C1, N, M);
    }
    // gu_T = rowsum(G); q_T = colsum(G) b_T / nu
    skb_rowmv_kernel<RM_RAW><<<rgrid, 256, 0, st>>>(G, nullptr, 0, nullptr, 0, nullptr, gu0, R1, R1, C1, N, M);
    skb_colmv_kernel<CM_Q0><<<cgrid, cblock, 0, st>>>(G, nullptr, 0, hb + (size_t)T * C1, sb, hq + (size_t)(T - 1) * C1, sq, R1, C1, N, M);
    for (int t = T; t >= 1; --t) {
        // p_t = (gu - a_t (E q_t)) a_t / mu, gu = rowsum(G) for t = T and 0 otherwise
        skb_rowmv_kernel<RM_BWD><<<rgrid, 256, 0, st>>>(E, hq + (size_t)(t - 1) * C1, sq, ha + (size_t)(t - 1) * R1, sa, t == T ? gu0 : nullptr,
                                                       hp + (size_t)(t - 1) * R1, sa, R1, C1, N, M);
        // q_{t-1} = -b_{t-1} (E^T p_t) b_{t-1} / nu  (gv for v_0 is not needed: v_0 = 0 is a constant)
        if (t > 1)
            skb_colmv_kernel<CM_BWD><<<cgrid, cblock, 0, st>>>(E, hp + (size_t)(t - 1) * R1, sa, hb + (size_t)(t - 1) * C1, sb,
                                                              hq + (size_t)(t - 2) * C1, sq, R1, C1, N, M);
    }
    const dim3 fgrid((C1 + 63) / 64, (R1 + 63) / 64, B);
    skb_final_kernel<<<fgrid, 256, 0, st>>>(G, E, ha, hp, hb, hq, gC, R1, C1, T);
    count_launch(4 * T + 4);
    return cudaGetLastError();
}

}  // namespace mdgat
