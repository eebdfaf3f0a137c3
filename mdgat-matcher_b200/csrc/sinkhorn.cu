// Log-domain Sinkhorn optimal transport (log_optimal_transport / log_sinkhorn_iterations,
// /root/reference/models/mdgat.py:279-308), float64, couplings resident in global memory / L2.
//
// The reference materialises Z + v (and Z + u) every half-iteration and then runs a
// max / sub / exp / sum / log chain over it; here a half-iteration is one kernel that reads
// the couplings and the opposite potential and writes one potential:
//     u_i = log_mu_i - LSE_j(C_ij + v_j)          (row pass, one warp per row)
//     v_j = log_nu_j - LSE_i(C_ij + u_i)          (column pass, 32 columns x 16 row-groups per CTA)
// The assignment matrix Z = C + u + v - norm is never written unless a caller asks for it.
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

__global__ void fill_dustbin_kernel(double* __restrict__ C, const double* __restrict__ bin_score, int N, int M) {
    // dustbin row i = N, column j = M and the corner = alpha (mdgat.py:294-299)
    const int b = blockIdx.y;
    double* Cb = C + (long long)b * (N + 1) * (M + 1);
    const double a = *bin_score;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= M) Cb[(long long)N * (M + 1) + t] = a;
    if (t < N) Cb[(long long)t * (M + 1) + M] = a;
}

cudaError_t launch_fill_dustbin(double* C, const double* bin_score, int B, int N, int M, cudaStream_t st) {
    const int n = max(N, M + 1);
    dim3 grid((n + 255) / 256, B);
    fill_dustbin_kernel<<<grid, 256, 0, st>>>(C, bin_score, N, M);
    count_launch();
    return cudaGetLastError();
}

// rows: R1 = N+1, cols: C1 = M+1.  first != 0: v is all zeros (iteration 0).
__global__ void __launch_bounds__(256)
sinkhorn_row_kernel(const double* __restrict__ C, const double* __restrict__ v, double* __restrict__ u,
                    int N, int M, double norm, int first) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int R1 = N + 1, C1 = M + 1;
    const int b = blockIdx.y;
    const int i = blockIdx.x * 8 + warp;
    if (i >= R1) return;
    const double* row = C + ((long long)b * R1 + i) * C1;
    const double* vb = v + (long long)b * C1;
    double mx = -INFINITY;
    for (int j = lane; j < C1; j += 32) mx = fmax(mx, row[j] + (first ? 0.0 : vb[j]));
    mx = warp_max_d(mx);
    double sum = 0.0;
    for (int j = lane; j < C1; j += 32) sum += exp(row[j] + (first ? 0.0 : vb[j]) - mx);
    sum = warp_sum_d(sum);
    if (lane == 0) {
        const double log_mu = (i < N) ? norm : (log((double)M) + norm);
        u[(long long)b * R1 + i] = log_mu - (mx + log(sum));
    }
}

constexpr int SK_TY = 16;
__global__ void __launch_bounds__(32 * SK_TY)
sinkhorn_col_kernel(const double* __restrict__ C, const double* __restrict__ u, double* __restrict__ v,
                    int N, int M, double norm) {
    __shared__ double red[SK_TY][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int R1 = N + 1, C1 = M + 1;
    const int b = blockIdx.y;
    const int j = blockIdx.x * 32 + tx;
    const double* Cb = C + (long long)b * R1 * C1;
    const double* ub = u + (long long)b * R1;
    const bool ok = j < C1;
    double mx = -INFINITY;
    if (ok) for (int i = ty; i < R1; i += SK_TY) mx = fmax(mx, Cb[(long long)i * C1 + j] + ub[i]);
    red[ty][tx] = mx;
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int t = 1; t < SK_TY; ++t) mx = fmax(mx, red[t][tx]);
        red[0][tx] = mx;
    }
    __syncthreads();
    mx = red[0][tx];
    __syncthreads();
    double sum = 0.0;
    if (ok) for (int i = ty; i < R1; i += SK_TY) sum += exp(Cb[(long long)i * C1 + j] + ub[i] - mx);
    red[ty][tx] = sum;
    __syncthreads();
    if (ty == 0 && ok) {
#pragma unroll
        for (int t = 1; t < SK_TY; ++t) sum += red[t][tx];
        const double log_nu = (j < M) ? norm : (log((double)N) + norm);
        v[(long long)b * C1 + j] = log_nu - (mx + log(sum));
    }
}

cudaError_t launch_sinkhorn(const double* C, double* u, double* v, int B, int N, int M, int iters, cudaStream_t st) {
    const double norm = -log((double)(N + M));
    dim3 rgrid((N + 1 + 7) / 8, B), cgrid((M + 1 + 31) / 32, B), cblock(32, SK_TY);
    if (iters <= 0) {
        cudaError_t e = cudaMemsetAsync(u, 0, sizeof(double) * (size_t)B * (N + 1), st);
        if (e != cudaSuccess) return e;
        return cudaMemsetAsync(v, 0, sizeof(double) * (size_t)B * (M + 1), st);
    }
    for (int it = 0; it < iters; ++it) {
        sinkhorn_row_kernel<<<rgrid, 256, 0, st>>>(C, v, u, N, M, norm, it == 0);
        sinkhorn_col_kernel<<<cgrid, cblock, 0, st>>>(C, u, v, N, M, norm);
    }
    count_launch(2 * iters);
    return cudaGetLastError();
}

}  // namespace mdgat
