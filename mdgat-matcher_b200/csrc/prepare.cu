// Input side of the path: what SparseDataset.__getitem__
// (/root/reference/load_data.py:213-292) computes per item on the CPU before the network runs,
// batched on the device:
//   * keypoints to world coordinates, kp_w = pose * T_cam0_velo * kp            (load_data.py:241-245)
//   * ground-truth matches from nearest neighbours under a distance threshold,
//     with or without the mutual check                                           (load_data.py:257-285)
//   * T_gt = inv(T_cam0_velo) inv(pose1) pose2 T_cam0_velo                       (load_data.py:238)
//   * "repeatability" count                                                      (load_data.py:266)
// One CTA per pair; world coordinates live in shared memory (N + M <= 4096 points).
#include "common.cuh"
#include "kernels.h"

namespace mdgat {

constexpr int PP_THREADS = 256;

DEVINL void mat4_mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            C[i * 4 + j] = s;
        }
}

// general 4x4 inverse by Gauss-Jordan with partial pivoting (poses are rigid, calib is affine)
__device__ void mat4_inv(const double* A, double* Inv) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { a[i][j] = A[i * 4 + j]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 8; ++j) { const double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        const double d = 1.0 / a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] *= d;
        for (int r = 0; r < 4; ++r) if (r != c) {
            const double f = a[r][c];
            for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) Inv[i * 4 + j] = a[i][4 + j];
}

__global__ void __launch_bounds__(PP_THREADS)
prepare_pairs_kernel(const double* __restrict__ kp1, const double* __restrict__ kp2,
                     const double* __restrict__ pose1, const double* __restrict__ pose2,
                     const double* __restrict__ T_cam0_velo, int calib_per_pair,
                     int N, int M, double threshold, int mutual_check,
                     int16_t* __restrict__ match1, int16_t* __restrict__ match2,
                     double* __restrict__ T_gt, int* __restrict__ rep) {
    extern __shared__ __align__(16) double sm[];
    double* w1 = sm;                 // [N][3] world coordinates of set 1
    double* w2 = w1 + 3 * N;         // [M][3]
    int* nn_of_row = reinterpret_cast<int*>(w2 + 3 * M);   // [N]  argmin over columns (min2)
    int* nn_of_col = nn_of_row + N;                        // [M]  argmin over rows (min1)
    double* dmin_row = reinterpret_cast<double*>(nn_of_col + M + ((N + M) & 1));   // [N] min1v
    double* dmin_col = dmin_row + N;                       // [M] min2v
    __shared__ double A1[16], A2[16];
    __shared__ int rep_s;
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* calib = T_cam0_velo + (calib_per_pair ? (long long)b * 16 : 0);
    if (tid == 0) {
        mat4_mul(pose1 + (long long)b * 16, calib, A1);
        mat4_mul(pose2 + (long long)b * 16, calib, A2);
        double ic[16], ip[16], t0[16], t1[16];
        mat4_inv(calib, ic);
        mat4_inv(pose1 + (long long)b * 16, ip);
        mat4_mul(ic, ip, t0);
        mat4_mul(t0, pose2 + (long long)b * 16, t1);
        mat4_mul(t1, calib, T_gt + (long long)b * 16);
        rep_s = 0;
    }
    __syncthreads();
    for (int i = tid; i < N + M; i += PP_THREADS) {
        const bool first = i < N;
        const double* p = first ? kp1 + ((long long)b * N + i) * 3 : kp2 + ((long long)b * M + (i - N)) * 3;
        const double* A = first ? A1 : A2;
        double* o = first ? w1 + 3 * i : w2 + 3 * (i - N);
#pragma unroll
        for (int r = 0; r < 3; ++r) o[r] = A[r * 4 + 0] * p[0] + A[r * 4 + 1] * p[1] + A[r * 4 + 2] * p[2] + A[r * 4 + 3];
    }
    __syncthreads();
    // nearest neighbour of every row (over columns) and of every column (over rows); first index on ties
    for (int i = tid; i < N + M; i += PP_THREADS) {
        const bool row = i < N;
        const double* me = row ? w1 + 3 * i : w2 + 3 * (i - N);
        const double* other = row ? w2 : w1;
        const int cnt = row ? M : N;
        double best = INFINITY; int bi = 0;
        for (int j = 0; j < cnt; ++j) {
            const double dx = me[0] - other[3 * j], dy = me[1] - other[3 * j + 1], dz = me[2] - other[3 * j + 2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best) { best = d2; bi = j; }
        }
        if (row) { nn_of_row[i] = bi; dmin_row[i] = sqrt(best); }
        else { nn_of_col[i - N] = bi; dmin_col[i - N] = sqrt(best); }
    }
    __syncthreads();
    int16_t* m1 = match1 + (long long)b * N;
    int16_t* m2 = match2 + (long long)b * M;
    int local_rep = 0;
    for (int i = tid; i < N; i += PP_THREADS) {
        const bool ok = dmin_row[i] < threshold;
        local_rep += ok ? 1 : 0;
        if (!mutual_check) m1[i] = ok ? (int16_t)nn_of_row[i] : (int16_t)-1;
        else m1[i] = -1;
    }
    for (int j = tid; j < M; j += PP_THREADS) {
        if (!mutual_check) m2[j] = dmin_col[j] < threshold ? (int16_t)nn_of_col[j] : (int16_t)-1;
        else m2[j] = -1;
    }
    if (local_rep) atomicAdd(&rep_s, local_rep);
    __syncthreads();
    if (mutual_check) {
        // column j is kept iff its nearest row i = min1[j] has j as nearest column (min2[i] == j) and row i is
        // within the threshold (j in min1f); load_data.py:275-280
        for (int j = tid; j < M; j += PP_THREADS) {
            const int i = nn_of_col[j];
            if (nn_of_row[i] == j && dmin_row[i] < threshold) { m1[i] = (int16_t)j; m2[j] = (int16_t)i; }
        }
    }
    if (tid == 0) rep[b] = rep_s;
}

size_t prepare_pairs_smem(int N, int M) {
    return (size_t)(3 * (N + M)) * sizeof(double) + (size_t)(N + M + ((N + M) & 1)) * sizeof(int) + (size_t)(N + M) * sizeof(double);
}

cudaError_t launch_prepare_pairs(const double* kp1, const double* kp2, const double* pose1, const double* pose2,
                                 const double* calib, int calib_per_pair, int B, int N, int M, double threshold,
                                 int mutual_check, int16_t* match1, int16_t* match2, double* T_gt, int* rep,
                                 cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    const size_t smem = prepare_pairs_smem(N, M);
    cudaError_t e = cudaFuncSetAttribute(prepare_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    prepare_pairs_kernel<<<B, PP_THREADS, smem, st>>>(kp1, kp2, pose1, pose2, calib, calib_per_pair, N, M, threshold,
                                                     mutual_check, match1, match2, T_gt, rep);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mdgat
