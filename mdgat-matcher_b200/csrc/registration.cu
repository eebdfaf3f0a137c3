// Batched one-shot rigid registration and match metrics on the device -- the post-processing the
// reference does pair by pair on the CPU after a D2H copy:
//   solve_icp()        /root/reference/utils/utils_test.py:73-110  (Kabsch: centroids, 3x3 SVD, R = U V^T)
//   calculate_error2() /root/reference/utils/utils_test.py:27-39   (RTE, RRE against T_gt)
//   match statistics   /root/reference/test_registration_metric.py:216-246 (TP / FP / TN / FN counts)
// One CTA per pair. Like the reference, R = U V^T is used as is (no determinant correction).
#include "common.cuh"
#include "kernels.h"
#include "../../include/mdgat_b200.h"

namespace mdgat {

constexpr int RG_THREADS = 256, RG_WARPS = RG_THREADS / 32;

DEVINL double ld_f64(const void* p, long long i, int dtype) {
    return dtype == MDGAT_F64 ? reinterpret_cast<const double*>(p)[i] : (double)reinterpret_cast<const float*>(p)[i];
}

// deterministic CTA-wide sums of NV values per thread; result valid in every thread
template <int NV>
DEVINL void block_sum(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum_d(v[k]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) red[warp * NV + k] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < RG_WARPS; ++w) s += red[w * NV + k];
        v[k] = s;
    }
}

// One-sided Jacobi SVD of a 3x3 matrix: A V = U diag(s). Returns R = U V^T.
__device__ void kabsch_rotation(const double H[3][3], double R[3][3]) {
    double A[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { A[i][j] = H[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) { alpha += A[k][p] * A[k][p]; beta += A[k][q] * A[k][q]; gamma += A[k][p] * A[k][q]; }
            const double lim = 1e-17 * sqrt(alpha * beta);
            off = fmax(off, fabs(gamma) - lim);
            if (fabs(gamma) > lim && gamma != 0.0) {
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double ap = A[k][p], aq = A[k][q];
                    A[k][p] = c * ap - s * aq; A[k][q] = s * ap + c * aq;
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
                }
            }
        }
        if (off <= 0.0) break;
    }
    double U[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double n = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
        const double inv = n > 0.0 ? 1.0 / n : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) U[k][j] = A[k][j] * inv;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + U[i][2] * V[j][2];
}

// stats per pair: [n_valid, n_valid_gt, tp, fp, tn, fn, rte, rre]
__global__ void __launch_bounds__(RG_THREADS)
register_pairs_kernel(const void* __restrict__ kpts0, const void* __restrict__ kpts1, int kp_dtype,
                      const int64_t* __restrict__ matches0, const int16_t* __restrict__ gt0,
                      const double* __restrict__ T_gt, int N, int M,
                      double* __restrict__ T_out, double* __restrict__ stats) {
    __shared__ double red[RG_WARPS * 9];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int64_t* m0 = matches0 + (long long)b * N;
    const long long k0 = (long long)b * N * 3, k1 = (long long)b * M * 3;

    // centroids of the matched keypoints (mkpts0 = target Q, mkpts1 = source P) and match counts
    double acc[7 + 5] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < N; i += RG_THREADS) {
        const long long m = m0[i];
        const bool valid = m > -1;
        if (valid) {
            acc[0] += 1.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                acc[1 + c] += ld_f64(kpts0, k0 + (long long)i * 3 + c, kp_dtype);
                acc[4 + c] += ld_f64(kpts1, k1 + m * 3 + c, kp_dtype);
            }
        }
        if (gt0) {
            long long g = gt0[(long long)b * N + i];
            if (g == M) g = -1;                                      // test_registration_metric.py:225
            acc[7] += g > -1 ? 1.0 : 0.0;                            // valid_gt
            acc[8] += (valid && m == g) ? 1.0 : 0.0;                 // true positive
            acc[9] += (valid && m != g) ? 1.0 : 0.0;                 // false positive
            acc[10] += (!valid && g == -1) ? 1.0 : 0.0;              // true negative
            acc[11] += (!valid && g > -1) ? 1.0 : 0.0;               // false negative
        }
    }
    {
        double a9[9];
#pragma unroll
        for (int k = 0; k < 7; ++k) a9[k] = acc[k];
        a9[7] = a9[8] = 0.0;
        block_sum<9>(a9, red);
#pragma unroll
        for (int k = 0; k < 7; ++k) acc[k] = a9[k];
        double c5[9] = {acc[7], acc[8], acc[9], acc[10], acc[11], 0, 0, 0, 0};
        block_sum<9>(c5, red);
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[7 + k] = c5[k];
    }
    const double cnt = acc[0];
    double uq[3], up[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { uq[c] = cnt > 0 ? acc[1 + c] / cnt : 0.0; up[c] = cnt > 0 ? acc[4 + c] / cnt : 0.0; }

    // H = Q_centered^T P_centered (utils_test.py:100)
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < N; i += RG_THREADS) {
        const long long m = m0[i];
        if (m > -1) {
            double q[3], p[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                q[c] = ld_f64(kpts0, k0 + (long long)i * 3 + c, kp_dtype) - uq[c];
                p[c] = ld_f64(kpts1, k1 + m * 3 + c, kp_dtype) - up[c];
            }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) h[a * 3 + c] += q[a] * p[c];
        }
    }
    block_sum<9>(h, red);
    if (tid != 0) return;

    double H[3][3], R[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) H[a][c] = h[a * 3 + c];
    kabsch_rotation(H, R);
    double t[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) t[a] = uq[a] - (R[a][0] * up[0] + R[a][1] * up[1] + R[a][2] * up[2]);
    double* T = T_out + (long long)b * 16;
#pragma unroll
    for (int a = 0; a < 3; ++a) { T[a * 4 + 0] = R[a][0]; T[a * 4 + 1] = R[a][1]; T[a * 4 + 2] = R[a][2]; T[a * 4 + 3] = t[a]; }
    T[12] = 0.0; T[13] = 0.0; T[14] = 0.0; T[15] = 1.0;

    double* st = stats + (long long)b * 8;
    st[0] = cnt; st[1] = acc[7]; st[2] = acc[8]; st[3] = acc[9]; st[4] = acc[10]; st[5] = acc[11];
    double rte = nan(""), rre = nan("");
    if (T_gt) {
        // T_error = inv(T) T_gt with inv(T) = [R^T | -R^T t] (R = U V^T is orthogonal); utils_test.py:34-38
        const double* G = T_gt + (long long)b * 16;
        double tr = 0.0, e[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            tr += R[0][a] * G[0 * 4 + a] + R[1][a] * G[1 * 4 + a] + R[2][a] * G[2 * 4 + a];
            e[a] = R[0][a] * (G[3] - t[0]) + R[1][a] * (G[7] - t[1]) + R[2][a] * (G[11] - t[2]);
        }
        rte = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        rre = acos((tr - 1.0) / 2.0);                              // NaN outside [-1, 1], like np.arccos
    }
    st[6] = rte; st[7] = rre;
}

cudaError_t launch_register_pairs(const void* kpts0, const void* kpts1, int kp_dtype, const int64_t* matches0,
                                  const int16_t* gt0, const double* T_gt, int B, int N, int M,
                                  double* T_out, double* stats, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    register_pairs_kernel<<<B, RG_THREADS, 0, st>>>(kpts0, kpts1, kp_dtype, matches0, gt0, T_gt, N, M, T_out, stats);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mdgat
