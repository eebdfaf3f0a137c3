// Match extraction and the three losses, straight from (couplings, u, v) without writing Z.
// Restates /root/reference/models/mdgat.py:442-483 (both the dustbin-arg-max variant used by
// loss_method != 'superglue' and the thresholded variant), :512-546 (triplet_loss), :547-594 (gap_loss) and
// :487-511 ('superglue' loss).
//
// Z_ij = ((C_ij + u_i) + v_j) - norm is evaluated in the reference's association order so
// that exact ties (duplicated keypoints, load_data.py:198-201) tie here as well; arg-max
// and top-2 take the lowest index among equals, like torch.max / the stable oracle.
#include "common.cuh"
#include "kernels.h"
#include "../../include/mdgat_b200.h"

namespace mdgat {

struct Top2 { double v1, v2; int i1, i2; };

DEVINL bool better(double av, int ai, double bv, int bi) { return av > bv || (av == bv && ai < bi); }
DEVINL void top2_insert(Top2& t, double v, int i) {
    if (better(v, i, t.v1, t.i1)) { t.v2 = t.v1; t.i2 = t.i1; t.v1 = v; t.i1 = i; }
    else if (better(v, i, t.v2, t.i2)) { t.v2 = v; t.i2 = i; }
}
DEVINL void top2_init(Top2& t) { t.v1 = t.v2 = -INFINITY; t.i1 = t.i2 = 0x7fffffff; }

// one warp per row i < N
__global__ void __launch_bounds__(256)
match_rows_kernel(const double* __restrict__ C, const double* __restrict__ u, const double* __restrict__ v,
                  int N, int M, int jlim, double norm,
                  double* __restrict__ rv1, double* __restrict__ rv2, int* __restrict__ ri1, int* __restrict__ ri2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, i = blockIdx.x * 8 + warp;
    if (i >= N) return;
    const int C1 = M + 1;
    const double* row = C + ((long long)b * (N + 1) + i) * C1;
    const double* vb = v + (long long)b * C1;
    const double ui = u[(long long)b * (N + 1) + i];
    Top2 t; top2_init(t);
    for (int j = lane; j < jlim; j += 32) top2_insert(t, ((row[j] + ui) + vb[j]) - norm, j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov1 = shfl_xor_d(t.v1, o), ov2 = shfl_xor_d(t.v2, o);
        const int oi1 = __shfl_xor_sync(0xffffffffu, t.i1, o), oi2 = __shfl_xor_sync(0xffffffffu, t.i2, o);
        top2_insert(t, ov1, oi1);
        top2_insert(t, ov2, oi2);
    }
    if (lane == 0) {
        const long long r = (long long)b * N + i;
        rv1[r] = t.v1; rv2[r] = t.v2; ri1[r] = t.i1; ri2[r] = t.i2;
    }
}

constexpr int MT_TY = 16;
__global__ void __launch_bounds__(32 * MT_TY)
match_cols_kernel(const double* __restrict__ C, const double* __restrict__ u, const double* __restrict__ v,
                  int N, int M, int ilim, double norm,
                  double* __restrict__ cv1, double* __restrict__ cv2, int* __restrict__ ci1, int* __restrict__ ci2) {
    __shared__ double sv1[MT_TY][33], sv2[MT_TY][33];
    __shared__ int si1[MT_TY][33], si2[MT_TY][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.y, j = blockIdx.x * 32 + tx;
    const int C1 = M + 1;
    const double* Cb = C + (long long)b * (N + 1) * C1;
    const double* ub = u + (long long)b * (N + 1);
    Top2 t; top2_init(t);
    if (j < M) {
        const double vj = v[(long long)b * C1 + j];
        for (int i = ty; i < ilim; i += MT_TY) top2_insert(t, ((Cb[(long long)i * C1 + j] + ub[i]) + vj) - norm, i);
    }
    sv1[ty][tx] = t.v1; sv2[ty][tx] = t.v2; si1[ty][tx] = t.i1; si2[ty][tx] = t.i2;
    __syncthreads();
    if (ty == 0 && j < M) {
#pragma unroll
        for (int k = 1; k < MT_TY; ++k) { top2_insert(t, sv1[k][tx], si1[k][tx]); top2_insert(t, sv2[k][tx], si2[k][tx]); }
        const long long c = (long long)b * M + j;
        cv1[c] = t.v1; cv2[c] = t.v2; ci1[c] = t.i1; ci2[c] = t.i2;
    }
}

DEVINL double neglogexp(double z) { return -log(exp(z)); }     // literally mdgat.py:541-542

__global__ void __launch_bounds__(256)
match_finalize_kernel(MatchParams p, double norm,
                      const double* __restrict__ rv1, const double* __restrict__ rv2,
                      const int* __restrict__ ri1, const int* __restrict__ ri2,
                      const double* __restrict__ cv1, const double* __restrict__ cv2,
                      const int* __restrict__ ci1, const int* __restrict__ ci2,
                      double* __restrict__ terms) {
    const int N = p.N, M = p.M, C1 = M + 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrow = (long long)p.B * N, ncol = (long long)p.B * M;
    if (t >= nrow + ncol) return;
    const bool thr_mode = p.match_mode == MDGAT_MATCH_THRESHOLD;
    if (p.bad != nullptr) {
        // A pair with a NaN / Inf input has an all-NaN assignment matrix in the reference: max() returns NaN with index 0 for
        // every row and column, so (mdgat.py:442-483) the dustbin variant reports match 0 everywhere with score NaN (with the
        // mutual check only row / column 0 is mutual, the other scores are 0), the threshold variant rejects everything
        // (NaN > threshold is false; with the mutual check score NaN at index 0), and every loss is NaN.
        const bool row = t < nrow;
        const long long c = row ? t : t - nrow;
        const int b = (int)(c / (row ? N : M)), i = (int)(c - (long long)b * (row ? N : M));
        if (p.bad[b]) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            int64_t* mo = row ? p.matches0 : p.matches1;
            double* so = row ? p.ms0 : p.ms1;
            mo[c] = thr_mode ? (int64_t)-1 : (int64_t)0;
            so[c] = thr_mode ? ((p.mutual_check && i == 0) ? nan : 0.0) : ((!p.mutual_check || i == 0) ? nan : 0.0);
            if (row && !thr_mode) atomicAdd(p.nvalid0, 1);
            if (p.loss_mode == MDGAT_LOSS_TRIPLET) terms[(long long)b * (N + M) + (row ? 0 : N) + i] = nan;
            return;
        }
    }
    if (t < nrow) {
        const int b = (int)(t / N), i = (int)(t - (long long)b * N);
        const int idx = ri1[t];
        const double mx = rv1[t];
        bool valid; double ms;
        if (!thr_mode) {
            valid = idx < M;
            bool keep = valid;
            if (p.mutual_check && valid) keep = (ci1[(long long)b * M + idx] == i);
            ms = keep ? exp(mx) : 0.0;
        } else {
            const double e = exp(mx);
            if (p.mutual_check) {
                const bool mutual = (ci1[(long long)b * M + idx] == i);
                ms = mutual ? e : 0.0;
                valid = mutual && (ms > p.match_threshold);
            } else {
                valid = e > p.match_threshold;
                ms = valid ? e : 0.0;
            }
        }
        p.matches0[t] = valid ? (int64_t)idx : (int64_t)-1;
        p.ms0[t] = ms;
        if (valid) atomicAdd(p.nvalid0, 1);
        if (p.loss_mode == MDGAT_LOSS_TRIPLET) {
            int pos = p.gt0[t]; if (pos < 0) pos = M;                      // mdgat.py:519
            const bool hit = (idx == pos);
            const double zneg = hit ? rv2[t] : rv1[t];                     // hard negative, :527-530
            const double* row = p.C + ((long long)b * (N + 1) + i) * C1;
            const double zpos = ((row[pos] + p.u[(long long)b * (N + 1) + i]) + p.v[(long long)b * C1 + pos]) - norm;
            terms[(long long)b * (N + M) + i] = fmax(neglogexp(zpos) - neglogexp(zneg) + p.gamma, 0.0);
        }
    } else {
        const long long c = t - nrow;
        const int b = (int)(c / M), j = (int)(c - (long long)b * M);
        const int idx = ci1[c];
        const double mx = cv1[c];
        bool valid; double ms;
        if (!thr_mode) {
            valid = idx < N;
            bool keep = valid;
            if (p.mutual_check && valid) keep = (ri1[(long long)b * N + idx] == j);
            ms = keep ? exp(mx) : 0.0;
        } else {
            const double e = exp(mx);
            if (p.mutual_check) {
                const bool mutual1 = (ri1[(long long)b * N + idx] == j);
                // mscores0 / valid0 of row idx (mdgat.py:450-453)
                const long long r = (long long)b * N + idx;
                const bool mutual0 = (ci1[(long long)b * M + ri1[r]] == idx);
                const double ms0 = mutual0 ? exp(rv1[r]) : 0.0;
                const bool valid0 = mutual0 && (ms0 > p.match_threshold);
                ms = mutual1 ? ms0 : 0.0;
                valid = mutual1 && valid0;
            } else {
                valid = e > p.match_threshold;
                ms = valid ? e : 0.0;
            }
        }
        p.matches1[c] = valid ? (int64_t)idx : (int64_t)-1;
        p.ms1[c] = ms;
        if (p.loss_mode == MDGAT_LOSS_TRIPLET) {
            int pos = p.gt1[c]; if (pos < 0) pos = N;                      // mdgat.py:520
            const bool hit = (idx == pos);
            const double zneg = hit ? cv2[c] : cv1[c];
            const double zpos = ((p.C[((long long)b * (N + 1) + pos) * C1 + j] + p.u[(long long)b * (N + 1) + pos])
                                 + p.v[(long long)b * C1 + j]) - norm;
            terms[(long long)b * (N + M) + N + j] = fmax(neglogexp(zpos) - neglogexp(zneg) + p.gamma, 0.0);
        }
    }
}

// deterministic mean of n terms (fixed reduction tree -> run-to-run identical loss)
__global__ void __launch_bounds__(1024)
mean_kernel(const double* __restrict__ terms, long long n, double* __restrict__ out) {
    __shared__ double red[1024];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) s += terms[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0] / (double)n;
}

// ------------------------------------------------------------------------------------------
// gap_loss (mdgat.py:547-594) and the 'superglue' loss (mdgat.py:487-511) straight from (C, u, v): Z is not written.
//
// Direction pc0 -> pc1 of gap_loss is row-aligned: row i has one positive (column gt0_i) and M negatives, and the
// reference's boolean-mask gathers keep their order. Direction pc1 -> pc0 is NOT: `scores[:,:,:-1][pos_match].view(b,m)`
// and `[neg_match].view(b,n,m)` flatten the (N+1) x M block ROW-major, so the M positives arrive sorted by (gt1_j, j) and the
// N M negatives are the non-positive entries in row-major order, cut into N rows of M (mdgat.py:580-584). Entry t of that
// list is the block entry e = t + #{k : p_k - k <= t}, p_0 < p_1 < ... the row-major positions of the positives. The kernels
// below reproduce exactly that pairing (it misaligns positives and negatives whenever gt1 is not sorted; the drop-in's
// contract is the reference's observable value, not the intended formula).
// clamp(x, min=0) of torch keeps NaN; fmax() would not.
// ------------------------------------------------------------------------------------------
DEVINL double clamp_min0(double x) { return x > 0.0 ? x : (x != x ? x : 0.0); }
DEVINL double z_at(const MatchParams& p, int b, int i, int j, double norm) {
    const int C1 = p.M + 1, R1 = p.N + 1;
    return ((p.C[((long long)b * R1 + i) * C1 + j] + p.u[(long long)b * R1 + i]) + p.v[(long long)b * C1 + j]) - norm;
}

// one warp per row i < N: rowterm = 2 log(sum_{j != gt0_i} max(0, nle(z_pos) - nle(z_ij) + gamma) + 1)
__global__ void __launch_bounds__(256)
gap_rows_kernel(MatchParams p, double norm, double* __restrict__ rowterm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, i = blockIdx.x * 8 + warp;
    if (i >= p.N) return;
    int pos = p.gt0[(long long)b * p.N + i]; if (pos < 0) pos = p.M;            // mdgat.py:554
    const double apos = neglogexp(z_at(p, b, i, pos, norm));
    double s = 0.0;
    for (int j = lane; j <= p.M; j += 32)
        if (j != pos) s += clamp_min0(apos - neglogexp(z_at(p, b, i, j, norm)) + p.gamma);
    s = warp_sum_d(s);
    if (lane == 0) rowterm[(long long)b * p.N + i] = 2.0 * log(s + 1.0);
}

// one CTA per pair: the positives of the (N+1) x M block in row-major order. q[k] = p_k - k, posval[k] = nle(z at p_k)
__global__ void __launch_bounds__(512)
gap_sort_kernel(MatchParams p, double norm, int* __restrict__ q, double* __restrict__ posval) {
    extern __shared__ int16_t g_gt[];
    const int b = blockIdx.x, N = p.N, M = p.M;
    for (int j = threadIdx.x; j < M; j += blockDim.x) { int g = p.gt1[(long long)b * M + j]; if (g < 0) g = N; g_gt[j] = (int16_t)g; }   // mdgat.py:555
    __syncthreads();
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const int g = g_gt[j];
        int rank = 0;
        for (int t = 0; t < M; ++t) { const int o = g_gt[t]; rank += (o < g || (o == g && t < j)) ? 1 : 0; }
        q[(long long)b * M + rank] = g * M + j - rank;
        posval[(long long)b * M + rank] = neglogexp(z_at(p, b, g, j, norm));
    }
}

// colterm[c] = 2 log(sum_r max(0, posval[c] - nle(neg[r][c]) + gamma) + 1), neg[r][c] = entry r M + c of the negative list
constexpr int GP_TY = 16;
__global__ void __launch_bounds__(32 * GP_TY)
gap_cols_kernel(MatchParams p, double norm, const int* __restrict__ q, const double* __restrict__ posval, double* __restrict__ colterm) {
    extern __shared__ int g_q[];                           // [M]
    __shared__ double red[GP_TY][33];
    const int tx = threadIdx.x, ty = threadIdx.y, N = p.N, M = p.M;
    const int b = blockIdx.y, c = blockIdx.x * 32 + tx;
    for (int t = ty * 32 + tx; t < M; t += 32 * GP_TY) g_q[t] = q[(long long)b * M + t];
    __syncthreads();
    double s = 0.0;
    if (c < M) {
        const double apos = posval[(long long)b * M + c];
        for (int r = ty; r < N; r += GP_TY) {
            const int t = r * M + c;
            int lo = 0, hi = M;                              // k = #{k : q_k <= t} (q is non-decreasing)
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (g_q[mid] <= t) lo = mid + 1; else hi = mid; }
            const int e = t + lo, i = e / M, j = e - i * M;
            s += clamp_min0(apos - neglogexp(z_at(p, b, i, j, norm)) + p.gamma);
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < M) {
#pragma unroll
        for (int k = 1; k < GP_TY; ++k) s += red[k][tx];
        colterm[(long long)b * M + c] = 2.0 * log(s + 1.0);
    }
}

// fixed-tree sum of n values by one CTA of 256 threads (every thread returns the total)
DEVINL double block_sum_256(const double* x, int n, double* red) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += x[i];
    __syncthreads();                                        // red may still be read from a previous call
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    return red[0];
}

// loss[b] = (mean_i rowterm + mean_c colterm) / 2 (mdgat.py:569, 592-594): one value per pair
__global__ void __launch_bounds__(256)
gap_final_kernel(const double* __restrict__ rowterm, const double* __restrict__ colterm, int N, int M, double* __restrict__ loss,
                 const int* __restrict__ bad) {
    __shared__ double red[256];
    const int b = blockIdx.x;
    if (bad != nullptr && bad[b]) { if (threadIdx.x == 0) loss[b] = __longlong_as_double(0x7ff8000000000000LL); return; }
    const double l0 = block_sum_256(rowterm + (long long)b * N, N, red) / (double)N;
    const double l1 = block_sum_256(colterm + (long long)b * M, M, red) / (double)M;
    if (threadIdx.x == 0) loss[b] = (l0 + l1) * 0.5;
}

// 'superglue' loss, per pair: (-sum_i z[i][gt0_i] - sum_{j: gt1_j = -1} z[N][j]) / (#{j: gt1_j = -1} + M); gt = -1 addresses the
// dustbin column / row as a negative index does in the reference (mdgat.py:493-509). N == M (the caller checks it).
__global__ void __launch_bounds__(256)
superglue_loss_kernel(MatchParams p, double norm, double* __restrict__ perpair) {
    __shared__ double red[256];
    __shared__ int cnt[256];
    const int b = blockIdx.x, N = p.N, M = p.M;
    double tp = 0.0, tn = 0.0;
    int xx = 0;
    for (int i = threadIdx.x; i < N; i += 256) {
        int g = p.gt0[(long long)b * N + i]; if (g < 0) g += M + 1;
        tp += z_at(p, b, i, g, norm);
    }
    for (int j = threadIdx.x; j < M; j += 256)
        if (p.gt1[(long long)b * M + j] == -1) { tn += z_at(p, b, N, j, norm); ++xx; }
    red[threadIdx.x] = tp; cnt[threadIdx.x] = xx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) { red[threadIdx.x] += red[threadIdx.x + o]; cnt[threadIdx.x] += cnt[threadIdx.x + o]; } __syncthreads(); }
    const double tps = red[0];
    const int xs = cnt[0];
    __syncthreads();
    red[threadIdx.x] = tn;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) perpair[b] = (p.bad != nullptr && p.bad[b]) ? __longlong_as_double(0x7ff8000000000000LL) : (-tps - red[0]) / (double)(xs + M);
}

__global__ void __launch_bounds__(256)
write_Z_kernel(const double* __restrict__ C, const double* __restrict__ u, const double* __restrict__ v,
               double* __restrict__ Z, int N, int M, double norm, long long total) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int C1 = M + 1, R1 = N + 1;
    const long long per = (long long)R1 * C1;
    const int b = (int)(t / per);
    const long long r = t - (long long)b * per;
    const int i = (int)(r / C1), j = (int)(r - (long long)i * C1);
    Z[t] = ((C[t] + u[(long long)b * R1 + i]) + v[(long long)b * C1 + j]) - norm;
}

cudaError_t launch_match_extract(const MatchParams& p, cudaStream_t st) {
    const int B = p.B, N = p.N, M = p.M;
    const double norm = -log((double)(N + M));
    const long long nrow = (long long)B * N, ncol = (long long)B * M;
    double* rv1 = p.scratch;
    double* rv2 = rv1 + nrow;
    double* cv1 = rv2 + nrow;
    double* cv2 = cv1 + ncol;
    double* terms = cv2 + ncol;
    int* ri1 = reinterpret_cast<int*>(terms + nrow + ncol);
    int* ri2 = ri1 + nrow;
    int* ci1 = ri2 + nrow;
    int* ci2 = ci1 + ncol;
    const bool thr = p.match_mode == MDGAT_MATCH_THRESHOLD;
    cudaError_t e = cudaMemsetAsync(p.nvalid0, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    dim3 rgrid((N + 7) / 8, B);
    match_rows_kernel<<<rgrid, 256, 0, st>>>(p.C, p.u, p.v, N, M, thr ? M : M + 1, norm, rv1, rv2, ri1, ri2);
    dim3 cgrid((M + 31) / 32, B), cblock(32, MT_TY);
    match_cols_kernel<<<cgrid, cblock, 0, st>>>(p.C, p.u, p.v, N, M, thr ? N : N + 1, norm, cv1, cv2, ci1, ci2);
    const long long tot = nrow + ncol;
    match_finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p, norm, rv1, rv2, ri1, ri2, cv1, cv2, ci1, ci2, terms);
    count_launch(3);
    if (p.loss_mode == MDGAT_LOSS_TRIPLET) { mean_kernel<<<1, 1024, 0, st>>>(terms, tot, p.loss); count_launch(); }
    if (p.loss_mode == MDGAT_LOSS_GAP) {
        // after match_finalize the second-best buffers are free: positives' order (q, posval) lives there
        int* q = ci2;
        double* posval = cv2;
        double* rowterm = terms;
        double* colterm = terms + nrow;
        gap_rows_kernel<<<rgrid, 256, 0, st>>>(p, norm, rowterm);
        gap_sort_kernel<<<B, 512, (size_t)M * sizeof(int16_t), st>>>(p, norm, q, posval);
        gap_cols_kernel<<<cgrid, dim3(32, GP_TY), (size_t)M * sizeof(int), st>>>(p, norm, q, posval, colterm);
        gap_final_kernel<<<B, 256, 0, st>>>(rowterm, colterm, N, M, p.loss, p.bad);
        count_launch(4);
    }
    if (p.loss_mode == MDGAT_LOSS_SUPERGLUE) {
        superglue_loss_kernel<<<B, 256, 0, st>>>(p, norm, terms);
        mean_kernel<<<1, 1024, 0, st>>>(terms, B, p.loss);
        count_launch(2);
    }
    if (p.Z) {
        const long long total = (long long)B * (N + 1) * (M + 1);
        write_Z_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.C, p.u, p.v, p.Z, N, M, norm, total);
        count_launch();
    }
    return cudaGetLastError();
}

}  // namespace mdgat
