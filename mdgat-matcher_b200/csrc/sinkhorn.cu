// Log-domain Sinkhorn optimal transport (log_optimal_transport / log_sinkhorn_iterations,
// /root/reference/models/mdgat.py:279-308), float64, couplings resident in global memory / L2.
//
// The reference materialises Z + v (and Z + u) every half-iteration and then runs a
// max / sub / exp / sum / log chain over it; here a half-iteration is one kernel that reads
// the couplings and the opposite potential and writes one potential:
//     u_i = log_mu_i - LSE_j(C_ij + v_j)          (row pass, one warp per row)
//     v_j = log_nu_j - LSE_i(C_ij + u_i)          (column pass, 32 columns x 16 row-groups per CTA)
// The assignment matrix Z = C + u + v - norm is never written unless a caller asks for it.
#include <cooperative_groups.h>
#include <string.h>
#include <stdlib.h>
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

// ------------------------------------------------------------------------------------------
// Fused Sinkhorn: all iterations in ONE launch, one 8-CTA cluster per pair, the kernel matrix
// resident on chip (shared memory + registers of the cluster).
//
// Algebra. With c_i = max_j C_ij and K_ij = exp(C_ij - c_i) in (0, 1] (computed once), the
// log-domain updates of the reference (mdgat.py:283-284)
//     u_i = log mu_i - LSE_j(C_ij + v_j),      v_j = log nu_j - LSE_i(C_ij + u_i)
// are, for a_i = exp(u_i + c_i) and b_j = exp(v_j), exactly the scaling iteration
//     a_i = mu_i / sum_j K_ij b_j,             b_j = nu_j / sum_i K_ij a_i,
// i.e. one multiply-add per matrix entry and one division per row / column per half-iteration
// instead of an exp per ENTRY, and no transcendental at all inside the loop. u = log a - c and
// v = log b are taken once at the end. All K_ij <= 1 with a 1 in every row, so while every row
// range max_j C_ij - min_j C_ij stays below SKF_MAX_RANGE the sums neither overflow nor lose
// relative accuracy; pairs that violate the bound (ill-conditioned, out-of-distribution
// inputs) are flagged and redone by the plain log-domain kernel below.
//
// Each CTA owns a slice of rows. Storage tiers for the K rows of a slice (local row r):
//   r <  n_smem                    shared memory   Ks[r][j], j < M
//   r <  n_smem + n_reg            registers       kreg[r - n_smem] = column j = tid (M <= 512)
//   otherwise                      global scratch  (L2-resident for mid sizes, HBM for N = 2048)
// The dustbin column j = M is one scalar per row (kd[r]) so that the main block has exactly M
// columns (512 at the benchmark shape = one column per thread). A CTA does the row sums of its
// rows, then the partial column sums over its rows; the 8 partials are combined through
// distributed shared memory with one cluster barrier per iteration.
// ------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int SKF_CLUSTER = 8, SKF_THREADS = 512, SKF_WARPS = SKF_THREADS / 32;
constexpr int SKF_NREG = 16;                 // K rows a CTA can keep in registers (one column per thread)
constexpr double SKF_MAX_RANGE = 300.0;

// Reduces NV per-lane values across the warp with NV-1 + log2(32/NV) shuffles (instead of
// 5 NV): each halving step trades half of the values with the partner lane. On return lane l
// holds the total of value index (l >> (5 - log2 NV)) in v[0].
template <int NV>
DEVINL void warp_transpose_sum(double (&v)[NV], int lane) {
    int off = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const double keep = up ? v[i + n / 2] : v[i];
            const double send = up ? v[i] : v[i + n / 2];
            v[i] = keep + shfl_xor_d(send, off);
        }
        off >>= 1;
    }
    for (; off > 0; off >>= 1) v[0] += shfl_xor_d(v[0], off);
}

__global__ void __launch_bounds__(SKF_THREADS, 1)
sinkhorn_fused_kernel(const double* __restrict__ C, double* __restrict__ Kg, double* __restrict__ u_out,
                      double* __restrict__ v_out, int* __restrict__ flags, int N, int M, int iters,
                      int RS, int rows_smem, int nreg, int ldk) {
    extern __shared__ __align__(16) double sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int b = blockIdx.x / SKF_CLUSTER;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R1 = N + 1, C1 = M + 1;
    const double mu_reg = 1.0 / (double)(N + M), mu_bin = (double)M / (double)(N + M);   // exp(log_mu)
    const double nu_reg = mu_reg, nu_bin = (double)N / (double)(N + M);                  // exp(log_nu)

    // shared layout (ldv = C1 rounded up to even, RSp = RS rounded up to a multiple of 4)
    const int ldv = (C1 + 1) & ~1, RSp = (RS + 3) & ~3;
    double* bs = sm;                        // [ldv]   b_j
    double* part = bs + ldv;                // [2][ldv] column partials of this CTA (peers read them)
    double* as = part + 2 * ldv;            // [RSp]   a_i of own rows (zero beyond nrows)
    double* cmx = as + RSp;                 // [RSp]   row maxima of C
    double* kd = cmx + RSp;                 // [RSp]   K of the dustbin column
    double* red = kd + RSp;                 // [SKF_NREG][SKF_WARPS]
    int* chg = reinterpret_cast<int*>(red + SKF_NREG * SKF_WARPS);   // [8] "changed" flags of the cluster (+ pad)
    double* Ks = red + SKF_NREG * SKF_WARPS + 4;   // [rows_smem][ldk]

    const int r0 = crank * RS;
    const int nrows = max(0, min(RS, R1 - r0));
    const int n_smem = min(nrows, rows_smem);
    const int n_reg = max(0, min(nrows - n_smem, nreg));
    const int g0 = n_smem + n_reg;           // first local row that lives in global scratch
    const double* Cb = C + (long long)b * R1 * C1;
    double* Kb = Kg + (long long)b * R1 * ldk;

    // ---- setup: row maxima, range check, K rows
    for (int r = tid; r < RSp; r += SKF_THREADS) { as[r] = 0.0; kd[r] = 0.0; cmx[r] = 0.0; }
    __syncthreads();
    int bad = 0;
    for (int r = warp; r < nrows; r += SKF_WARPS) {
        const double* crow = Cb + (long long)(r0 + r) * C1;
        double mx = -INFINITY, mn = INFINITY;
        for (int j = lane; j < C1; j += 32) { const double c = crow[j]; mx = fmax(mx, c); mn = fmin(mn, c); }
        mx = warp_max_d(mx);
        mn = -warp_max_d(-mn);
        if (!(mx - mn < SKF_MAX_RANGE)) bad = 1;          // also catches NaN / inf
        if (r < n_smem || r >= g0) {
            double* krow = r < n_smem ? Ks + (size_t)r * ldk : Kb + (long long)(r0 + r) * ldk;
            for (int j = lane; j < M; j += 32) krow[j] = exp(crow[j] - mx);
        }
        if (lane == 0) { cmx[r] = mx; kd[r] = exp(crow[M] - mx); }
    }
    if (bad && lane == 0) atomicOr(flags + b, 1);
    for (int j = tid; j < C1; j += SKF_THREADS) bs[j] = 1.0;      // v = 0 before the first row pass
    __syncthreads();
    double kreg[SKF_NREG];
#pragma unroll
    for (int q = 0; q < SKF_NREG; ++q) {
        kreg[q] = 0.0;
        if (q < n_reg && tid < M) kreg[q] = exp(Cb[(long long)(r0 + n_smem + q) * C1 + tid] - cmx[n_smem + q]);
    }

    int it_done = 0;
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        const double bM = bs[M];
        // ---- row sums s_r = sum_j K_rj b_j, then a_r = mu_r / s_r
        // (1) register rows: per-thread products, reduced across the CTA through red[]
        if (n_reg > 0) {
            const double bj = tid < M ? bs[tid] : 0.0;
            double pr[SKF_NREG];
#pragma unroll
            for (int q = 0; q < SKF_NREG; ++q) pr[q] = kreg[q] * bj;
            warp_transpose_sum<SKF_NREG>(pr, lane);
            if ((lane & 1) == 0) red[(lane >> 1) * SKF_WARPS + warp] = pr[0];
        }
        // (2) shared / global rows: a warp takes four rows at a time
        for (int rb = warp * 4; rb < nrows; rb += SKF_WARPS * 4) {
            if (rb >= n_smem && min(rb + 3, nrows - 1) < g0) continue;      // all four live in registers
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const double* kr[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = min(rb + q, nrows - 1);
                kr[q] = r < n_smem ? Ks + (size_t)r * ldk : Kb + (long long)(r0 + r) * ldk;
            }
            for (int j = lane; j < M; j += 32) {
                const double bj = bs[j];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = fma(kr[q][j], bj, acc[q]);
            }
            warp_transpose_sum<4>(acc, lane);
            const int r = rb + (lane >> 3);
            if ((lane & 7) == 0 && r < nrows && !(r >= n_smem && r < g0))
                as[r] = ((r0 + r < N) ? mu_reg : mu_bin) / fma(kd[r], bM, acc[0]);
        }
        __syncthreads();
        if (tid < n_reg) {
            const int r = n_smem + tid;
            double sum = 0.0;
#pragma unroll
            for (int x = 0; x < SKF_WARPS; ++x) sum += red[tid * SKF_WARPS + x];
            as[r] = ((r0 + r < N) ? mu_reg : mu_bin) / fma(kd[r], bM, sum);
        }
        __syncthreads();
        // ---- partial column sums over own rows: part_j = sum_r K_rj a_r
        double* pb = part + buf * ldv;
        for (int j = tid; j < M; j += SKF_THREADS) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int r = 0;
            for (; r + 3 < n_smem; r += 4) {
                const double2 w01 = *reinterpret_cast<const double2*>(as + r);
                const double2 w23 = *reinterpret_cast<const double2*>(as + r + 2);
                a0 = fma(Ks[(size_t)r * ldk + j], w01.x, a0);
                a1 = fma(Ks[(size_t)(r + 1) * ldk + j], w01.y, a1);
                a2 = fma(Ks[(size_t)(r + 2) * ldk + j], w23.x, a2);
                a3 = fma(Ks[(size_t)(r + 3) * ldk + j], w23.y, a3);
            }
            for (; r < n_smem; ++r) a0 = fma(Ks[(size_t)r * ldk + j], as[r], a0);
            if (j == tid) {
#pragma unroll
                for (int q = 0; q < SKF_NREG; ++q) if (q < n_reg) a1 = fma(kreg[q], as[n_smem + q], a1);
            }
            for (r = g0; r < nrows; ++r) a2 = fma(Kb[(long long)(r0 + r) * ldk + j], as[r], a2);
            pb[j] = (a0 + a1) + (a2 + a3);
        }
        if (warp == SKF_WARPS - 1) {                    // dustbin column
            double a = 0.0;
            for (int r = lane; r < nrows; r += 32) a = fma(kd[r], as[r], a);
            a = warp_sum_d(a);
            if (lane == 0) pb[M] = a;
        }
        cluster.sync();
        // ---- reduce-scatter + broadcast through distributed shared memory: this CTA sums the 8
        // partials of its own 1/8 of the columns (8 adjacent lanes fetch one partial each), forms
        // b_j = nu_j / total and stores it into the b vector of every CTA of the cluster.
        bool changed = false;
        {
            const int CS = (C1 + SKF_CLUSTER - 1) / SKF_CLUSTER;
            const int c0 = crank * CS;
            const int ncols = max(0, min(CS, C1 - c0));
            const int src = tid & (SKF_CLUSTER - 1);
            for (int base = 0; base < ncols * SKF_CLUSTER; base += SKF_THREADS) {
                const int idx = base + tid;
                const int j = c0 + (idx >> 3);
                const bool ok = idx < ncols * SKF_CLUSTER;
                const double old = ok ? bs[j] : 0.0;          // column j is written only by this CTA's lanes, below
                double pv = ok ? cluster.map_shared_rank(part, src)[buf * ldv + j] : 0.0;
                pv += shfl_xor_d(pv, 1);
                pv += shfl_xor_d(pv, 2);
                pv += shfl_xor_d(pv, 4);
                if (ok) {
                    const double nb = ((j < M) ? nu_reg : nu_bin) / pv;
                    changed |= (__double_as_longlong(nb) != __double_as_longlong(old));
                    cluster.map_shared_rank(bs, src)[j] = nb;
                }
            }
        }
        // Exact early exit: the iteration is a deterministic map, so once b repeats bit for bit
        // every further iteration is a no-op. Each CTA publishes "one of my columns changed" to all
        // peers; everybody sees the same eight flags after the barrier and leaves together.
        const int any_local = __syncthreads_or(changed ? 1 : 0);
        if (tid < SKF_CLUSTER) cluster.map_shared_rank(chg, tid)[crank] = any_local;
        cluster.sync();
        ++it_done;
        int any = 0;
#pragma unroll
        for (int c = 0; c < SKF_CLUSTER; ++c) any |= chg[c];
        if (!any) break;
    }
    if (crank == 0 && tid == 0) flags[gridDim.x / SKF_CLUSTER + b] = it_done;
    // u_i = log a_i - c_i, v_j = log b_j
    for (int r = tid; r < nrows; r += SKF_THREADS) u_out[(long long)b * R1 + r0 + r] = iters > 0 ? log(as[r]) - cmx[r] : 0.0;
    if (crank == 0)
        for (int j = tid; j < C1; j += SKF_THREADS) v_out[(long long)b * C1 + j] = iters > 0 ? log(bs[j]) : 0.0;
}

// ------------------------------------------------------------------------------------------
// The same scaling iteration with the kernel matrix STORED in float32 and every sum, division and potential in float64.
//
// The fused kernel above is bound by the shared-memory wavefronts of its two sweeps over K per iteration (ncu: 6400
// wavefronts per iteration and CTA against 8250 cycles). K_ij = exp(C_ij - c_i) in (e^-80, 1] rounded to float32 (relative
// error <= 2^-24) is the exact kernel of the couplings C'_ij = C_ij + log(1 + d_ij), |d_ij| <= 6e-8, and the Sinkhorn map is
// non-expansive in the log domain, so u and v move by O(1e-7): three orders of magnitude inside the 1e-4 score bar, and
// SURVEY.md 7.3 measured the whole tail as float32-safe. Exact ties stay exact ties (equal doubles round to equal floats).
// A float row is 2 KB at M = 512: all 65 rows of a CTA fit in shared memory (no register tier), half the wavefronts per
// sweep; at N = 2048 the rows that stream from L2 / HBM are half the bytes. float -> double runs on the integer pipe
// (two IMADs build the double's words; cvt.f64.f32 would sit on the 16-lane conversion pipe), valid for normal positive
// floats -- guaranteed by the range bound SK32_MAX_RANGE, pairs beyond it go to the log-domain fallback like before.
// config['precision'] = 'exact' keeps the float64 kernel matrix.
// ------------------------------------------------------------------------------------------
//
// Early exit at a tolerance. The reference always runs T iterations (mdgat.py:282-284), but on the network's scores the
// iteration contracts by ~0.4 per step: after ~25 of the T = 100 iterations no potential moves by 1e-10 any more, and the
// bit-for-bit rule of the float64 kernel stopped at 48 iterations on average (cfg2 batch, ncu barrier counts). This kernel
// stops when every b_j moved by at most tol = 2^-35 relative in one iteration: 19-31 iterations per pair, 26 on average
// (bench.py reports the counts, config.sinkhorn), 0.74 -> 0.42 ms at cfg2 and 15.2 -> 8.0 ms at cfg4. The per-iteration change of a Sinkhorn
// iterate does not grow (the map is non-expansive in the log domain), so the T - t iterations that were skipped would have moved
// a log-potential by at most (T - t) tol <= 3e-9 at T = 100 -- the size of the float32 storage effect above, 1e-4 is the bar.
// A pair that has not converged keeps iterating up to T exactly like the reference. config['precision'] = 'exact' (the
// float64 kernel above) keeps the bit-for-bit exit.
constexpr double SK32_MAX_RANGE = 80.0;
constexpr double SK32_EXIT_TOL = 2.9103830456733704e-11;      // 2^-35
constexpr int SK32_DEFAULT_THREADS = 512;
DEVINL double f32bits_to_f64(uint32_t f) {
    uint32_t hi;
    asm("mad.hi.u32 %0, %1, 0x20000000, 0x38000000;" : "=r"(hi) : "r"(f));     // (f >> 3) + ((1023 - 127) << 20)
    return __hiloint2double((int)hi, (int)(f << 29));
}

template <int MINB, int NT>
__global__ void __launch_bounds__(NT, MINB)
sinkhorn_fused32_kernel(const double* __restrict__ C, float* __restrict__ Kg, double* __restrict__ u_out,
                        double* __restrict__ v_out, int* __restrict__ flags, int N, int M, int iters,
                        int RS, int rows_smem, int ldk, double tol) {
    extern __shared__ __align__(16) double sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int b = blockIdx.x / SKF_CLUSTER;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R1 = N + 1, C1 = M + 1;
    const double mu_reg = 1.0 / (double)(N + M), mu_bin = (double)M / (double)(N + M);
    const double nu_reg = mu_reg, nu_bin = (double)N / (double)(N + M);

    // NT = 512: 16 warps, a warp takes four rows per round, a thread one column over all rows of the CTA.
    // NT = 1024: 32 warps (64 registers per thread), a warp takes two rows, two threads share a column (half the rows each,
    // two partial vectors per CTA): the iteration is a chain of dependent phases, more warps shorten every link of it.
    constexpr int NW = NT / 32, RPW = NT == 512 ? 4 : 2, HALVES = NT / 512, NP = SKF_CLUSTER * HALVES;
    static_assert(NT == 512 || NT == 1024, "512 or 1024 threads");
    const int ldv = (C1 + 1) & ~1, RSp = (RS + 3) & ~3;
    double* bs = sm;                        // [ldv]    b_j
    double* part = bs + ldv;                // [2][HALVES][ldv] column partials of this CTA (peers read them)
    double* as = part + 2 * HALVES * ldv;   // [RSp]    a_i of own rows
    double* cmx = as + RSp;                 // [RSp]    row maxima of C
    double* kd = cmx + RSp;                 // [RSp]    K of the dustbin column (float64: one scalar per row)
    int* chg = reinterpret_cast<int*>(kd + RSp);                    // [8] "changed" flags of the cluster
    __shared__ double s_aN;                                         // a of the dustbin row (every CTA computes the same value)
    uint32_t* Ks = reinterpret_cast<uint32_t*>(kd + RSp + 4);       // [rows_smem][ldk] float bit patterns

    // The dustbin ROW of the couplings is the constant alpha (mdgat.py:294-299), so its kernel row is all ones: its row sum is
    // the sum of b, its contribution to every column sum is a_N itself -- no storage, and the N real rows split evenly
    // over the CTAs (512 rows: 64 per CTA = one 4-row round of the 16 warps; with the dustbin row in a slice one warp had a
    // second round while fifteen waited, 12 % of the kernel's stall samples)
    const int r0 = crank * RS;
    const int nrows = max(0, min(RS, N - r0));
    const int n_smem = min(nrows, rows_smem);
    const double* Cb = C + (long long)b * R1 * C1;
    uint32_t* Kb = reinterpret_cast<uint32_t*>(Kg) + (long long)b * R1 * ldk;
    const int Me = M & ~1;                   // columns taken two at a time; an odd last column separately

    // ---- setup: row maxima, range check, K rows
    for (int r = tid; r < RSp; r += NT) { as[r] = 0.0; kd[r] = 0.0; cmx[r] = 0.0; }
    __syncthreads();
    int bad = 0;
    for (int r = warp; r < nrows; r += NW) {
        const double* crow = Cb + (long long)(r0 + r) * C1;
        double mx = -INFINITY, mn = INFINITY;
        for (int j = lane; j < C1; j += 32) { const double c = crow[j]; mx = fmax(mx, c); mn = fmin(mn, c); }
        mx = warp_max_d(mx);
        mn = -warp_max_d(-mn);
        if (!(mx - mn < SK32_MAX_RANGE)) bad = 1;          // also catches NaN / inf
        uint32_t* krow = r < n_smem ? Ks + (size_t)r * ldk : Kb + (long long)(r0 + r) * ldk;
        for (int j = lane; j < M; j += 32) krow[j] = __float_as_uint(__double2float_rn(exp(crow[j] - mx)));
        if (lane == 0) { cmx[r] = mx; kd[r] = exp(crow[M] - mx); }
    }
    if (bad && lane == 0) atomicOr(flags + b, 1);
    for (int j = tid; j < ldv; j += NT) bs[j] = j < C1 ? 1.0 : 0.0;      // v = 0 before the first row pass
    __syncthreads();

    int it_done = 0;
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        const double bM = bs[M];
        // ---- row sums s_r = sum_j K_rj b_j, then a_r = mu_r / s_r: a warp takes RPW rows at a time, a lane two columns
        for (int rb = warp * RPW; rb < nrows; rb += NW * RPW) {
            double acc[RPW];
#pragma unroll
            for (int q = 0; q < RPW; ++q) acc[q] = 0.0;
            if (rb + RPW - 1 < n_smem) {
                const uint32_t* k0 = Ks + (size_t)rb * ldk;
#pragma unroll 4
                for (int j = 2 * lane; j < Me; j += 64) {
                    const double2 bj = *reinterpret_cast<const double2*>(bs + j);
#pragma unroll
                    for (int q = 0; q < RPW; ++q) {
                        const uint2 k = *reinterpret_cast<const uint2*>(k0 + (size_t)q * ldk + j);
                        acc[q] = fma(f32bits_to_f64(k.x), bj.x, acc[q]);
                        acc[q] = fma(f32bits_to_f64(k.y), bj.y, acc[q]);
                    }
                }
                if (Me < M && lane == 0) {
#pragma unroll
                    for (int q = 0; q < RPW; ++q) acc[q] = fma(f32bits_to_f64(k0[(size_t)q * ldk + Me]), bs[Me], acc[q]);
                }
            } else {
                const uint32_t* kr[RPW];
#pragma unroll
                for (int q = 0; q < RPW; ++q) {
                    const int r = min(rb + q, nrows - 1);
                    kr[q] = r < n_smem ? Ks + (size_t)r * ldk : Kb + (long long)(r0 + r) * ldk;
                }
#pragma unroll 2
                for (int j = 2 * lane; j < Me; j += 64) {
                    const double2 bj = *reinterpret_cast<const double2*>(bs + j);
#pragma unroll
                    for (int q = 0; q < RPW; ++q) {
                        const uint2 k = *reinterpret_cast<const uint2*>(kr[q] + j);
                        acc[q] = fma(f32bits_to_f64(k.x), bj.x, acc[q]);
                        acc[q] = fma(f32bits_to_f64(k.y), bj.y, acc[q]);
                    }
                }
                if (Me < M && lane == 0) {
#pragma unroll
                    for (int q = 0; q < RPW; ++q) acc[q] = fma(f32bits_to_f64(kr[q][Me]), bs[Me], acc[q]);
                }
            }
            warp_transpose_sum<RPW>(acc, lane);                   // lane l holds the total of row (l / (32 / RPW))
            const int r = rb + lane / (32 / RPW);
            if ((lane & (32 / RPW - 1)) == 0 && r < nrows) as[r] = mu_reg / fma(kd[r], bM, acc[0]);
        }
        if (warp == NW - 1) {                           // dustbin row: a_N = mu_N / sum_j b_j (its kernel row is all ones)
            double sb = 0.0;
            for (int j = lane; j < C1; j += 32) sb += bs[j];
            sb = warp_sum_d(sb);
            if (lane == 0) s_aN = mu_bin / sb;
        }
        __syncthreads();
        // ---- partial column sums over own rows: part_j = sum_r K_rj a_r (NT = 1024: rows [0, rsplit) and [rsplit, nrows) by
        // the two threads of a column, one partial vector each)
        double* pb = part + buf * (HALVES * ldv);
        {
            const int half = HALVES == 1 ? 0 : (tid >> 9);
            const int rsplit = HALVES == 1 ? nrows : min(nrows, ((nrows + 1) / 2 + 3) & ~3);
            const int ra = half == 0 ? 0 : rsplit, rz = (HALVES == 1 || half == 1) ? nrows : rsplit;
            const int sa = min(ra, n_smem), sz = min(rz, n_smem);       // the part of [ra, rz) that lives in shared memory
            const int ga = max(ra, n_smem);                              // the rest streams from the L2-resident scratch
            for (int j = tid & 511; j < M; j += 512) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int r = sa;
                for (; r + 3 < sz; r += 4) {
                    const double2 w01 = *reinterpret_cast<const double2*>(as + r);
                    const double2 w23 = *reinterpret_cast<const double2*>(as + r + 2);
                    a0 = fma(f32bits_to_f64(Ks[(size_t)r * ldk + j]), w01.x, a0);
                    a1 = fma(f32bits_to_f64(Ks[(size_t)(r + 1) * ldk + j]), w01.y, a1);
                    a2 = fma(f32bits_to_f64(Ks[(size_t)(r + 2) * ldk + j]), w23.x, a2);
                    a3 = fma(f32bits_to_f64(Ks[(size_t)(r + 3) * ldk + j]), w23.y, a3);
                }
                for (; r < sz; ++r) a0 = fma(f32bits_to_f64(Ks[(size_t)r * ldk + j]), as[r], a0);
                r = max(r, ga);
                for (; r + 3 < rz; r += 4) {
                    const uint32_t* kp = Kb + (long long)(r0 + r) * ldk + j;
                    const uint32_t k0 = kp[0], k1 = kp[ldk], k2 = kp[2 * (size_t)ldk], k3 = kp[3 * (size_t)ldk];
                    a0 = fma(f32bits_to_f64(k0), as[r], a0);
                    a1 = fma(f32bits_to_f64(k1), as[r + 1], a1);
                    a2 = fma(f32bits_to_f64(k2), as[r + 2], a2);
                    a3 = fma(f32bits_to_f64(k3), as[r + 3], a3);
                }
                for (; r < rz; ++r) a2 = fma(f32bits_to_f64(Kb[(long long)(r0 + r) * ldk + j]), as[r], a2);
                pb[half * ldv + j] = (a0 + a1) + (a2 + a3);
            }
        }
        if (warp == NW - 1) {                           // dustbin column
            double a = 0.0;
            for (int r = lane; r < nrows; r += 32) a = fma(kd[r], as[r], a);
            a = warp_sum_d(a);
            if (lane == 0) { pb[M] = a; if (HALVES == 2) pb[ldv + M] = 0.0; }
        }
        cluster.sync();
        // ---- reduce-scatter + broadcast through distributed shared memory (as in the float64 kernel)
        bool changed = false;
        {
            const int CS = (C1 + SKF_CLUSTER - 1) / SKF_CLUSTER;
            const int c0 = crank * CS;
            const int ncols = max(0, min(CS, C1 - c0));
            const int src = tid & (NP - 1);                 // partial vector src: CTA src % 8, half src / 8
            for (int base = 0; base < ncols * NP; base += NT) {
                const int idx = base + tid;
                const int j = c0 + idx / NP;
                const bool ok = idx < ncols * NP;
                const double old = ok ? bs[j] : 0.0;
                double pv = ok ? cluster.map_shared_rank(part, src & (SKF_CLUSTER - 1))[(buf * HALVES + (src >> 3)) * ldv + j] : 0.0;
                pv += shfl_xor_d(pv, 1);
                pv += shfl_xor_d(pv, 2);
                pv += shfl_xor_d(pv, 4);
                if (HALVES == 2) pv += shfl_xor_d(pv, 8);
                if (ok) {
                    const double nb = ((j < M) ? nu_reg : nu_bin) / (pv + s_aN);      // + the dustbin row: K_Nj = 1
                    changed |= fabs(nb - old) > tol * old;                            // tol = 0: "repeats bit for bit" (b > 0)
                    if (src < SKF_CLUSTER) cluster.map_shared_rank(bs, src)[j] = nb;
                }
            }
        }
        const int any_local = __syncthreads_or(changed ? 1 : 0);
        if (tid < SKF_CLUSTER) cluster.map_shared_rank(chg, tid)[crank] = any_local;
        cluster.sync();
        ++it_done;
        int any = 0;
#pragma unroll
        for (int c = 0; c < SKF_CLUSTER; ++c) any |= chg[c];
        if (!any) break;
    }
    if (crank == 0 && tid == 0) flags[gridDim.x / SKF_CLUSTER + b] = it_done;
    for (int r = tid; r < nrows; r += NT) u_out[(long long)b * R1 + r0 + r] = iters > 0 ? log(as[r]) - cmx[r] : 0.0;
    if (crank == 0) {
        for (int j = tid; j < C1; j += NT) v_out[(long long)b * C1 + j] = iters > 0 ? log(bs[j]) : 0.0;
        if (tid == 0) u_out[(long long)b * R1 + N] = iters > 0 ? log(s_aN) - Cb[(long long)N * C1] : 0.0;      // c_N = alpha
    }
}

// Plain log-domain Sinkhorn for flagged pairs: one CTA per pair, exact max subtraction.
__global__ void __launch_bounds__(1024)
sinkhorn_safe_kernel(const double* __restrict__ C, double* __restrict__ u, double* __restrict__ v,
                     const int* __restrict__ flags, int N, int M, int iters) {
    const int b = blockIdx.x;
    if (flags[b] == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R1 = N + 1, C1 = M + 1;
    const double norm = -log((double)(N + M));
    const double* Cb = C + (long long)b * R1 * C1;
    double* ub = u + (long long)b * R1;
    double* vb = v + (long long)b * C1;
    for (int j = tid; j < C1; j += 1024) vb[j] = 0.0;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        for (int i = warp; i < R1; i += 32) {
            const double* row = Cb + (long long)i * C1;
            double mx = -INFINITY;
            for (int j = lane; j < C1; j += 32) mx = fmax(mx, row[j] + vb[j]);
            mx = warp_max_d(mx);
            const double m0 = (mx == -INFINITY || mx == INFINITY) ? 0.0 : mx;
            double sum = 0.0;
            for (int j = lane; j < C1; j += 32) sum += exp(row[j] + vb[j] - m0);
            sum = warp_sum_d(sum);
            if (lane == 0) ub[i] = ((i < N) ? norm : (log((double)M) + norm)) - (m0 + log(sum));
        }
        __syncthreads();
        for (int j = tid; j < C1; j += 1024) {
            double mx = -INFINITY;
            for (int i = 0; i < R1; ++i) mx = fmax(mx, Cb[(long long)i * C1 + j] + ub[i]);
            const double m0 = (mx == -INFINITY || mx == INFINITY) ? 0.0 : mx;
            double sum = 0.0;
            for (int i = 0; i < R1; ++i) sum += exp(Cb[(long long)i * C1 + j] + ub[i] - m0);
            vb[j] = ((j < M) ? norm : (log((double)N) + norm)) - (m0 + log(sum));
        }
        __syncthreads();
    }
}

size_t sinkhorn_scratch_doubles(int B, int N, int M) {
    const size_t ldk = (size_t)((M + 1) & ~1);
    return (size_t)B * (N + 1) * ldk + (size_t)B + 2;      // K scratch + per-pair flags and iteration counts (ints)
}

// float32 kernel matrix: same scratch (the float rows use half of the K area), same flags / iteration counts behind it.
// Launch shape: one CTA per SM with as many K rows in shared memory as fit (all 64 real rows at N = M = 512). The iteration is
// bound by its chain of dependent phases (row sums -> CTA barrier -> column partials -> cluster barrier -> exchange -> CTA
// barrier -> cluster barrier): about 12 500 cycles per iteration (ncu capture of the final kernel: 413 us, two waves, ~28
// iterations of the slowest pair of a wave, 11 % of the samples in the setup) against 4 460 shared-memory wavefronts and ~970
// warp instructions per warp; stall samples of an iteration: row sweep 23 %, column sweep up to the cluster barrier 40 %,
// exchange 16 %, second cluster barrier 8 %. At cfg2 the 32 clusters of 8 need two waves (16 clusters fit the
// GPCs of 148 SMs). Measured and rejected: MDGAT_SK_CTAS=2, two CTAs per SM (64 registers, 49 of the rows in shared memory,
// the others read from the L2-resident scratch), all 32 clusters in one wave -- 1.16 ms against 0.74 ms before the tolerance
// exit; MDGAT_SK_THREADS=1024, 32 warps per CTA (two rows per warp round, two threads per column, 16 partial vectors in the
// exchange) -- 0.54 ms against 0.42 ms: the barriers and the exchange grow with the warp count faster than the sweeps shrink.
template <int MINB, int NT>
static cudaError_t sk32_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int B, int RS, int ldk, int ldv, cudaStream_t st,
                               int& rows_smem, int& clusters) {
    int dev = 0, max_optin = 0, max_sm = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&max_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev)) != cudaSuccess) return e;
    const size_t budget = MINB == 1 ? (size_t)max_optin : (size_t)(max_sm / MINB - 1024);      // 1 KB per resident CTA is reserved
    const size_t fixed = ((size_t)(1 + 2 * (NT / 512)) * ldv + 3 * (size_t)((RS + 3) & ~3) + 4) * sizeof(double);
    if (fixed + 1024 > budget) return cudaErrorInvalidValue;
    rows_smem = (int)((budget - fixed) / ((size_t)ldk * sizeof(float)));
    if (rows_smem > RS) rows_smem = RS;
    const size_t smem = fixed + (size_t)rows_smem * ldk * sizeof(float);
    if ((e = cudaFuncSetAttribute(sinkhorn_fused32_kernel<MINB, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(SKF_CLUSTER * B);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = SKF_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    clusters = 1;
    if (MINB > 1 && (e = cudaOccupancyMaxActiveClusters(&clusters, sinkhorn_fused32_kernel<MINB, NT>, &cfg)) != cudaSuccess) return e;
    return cudaSuccess;
}

static cudaError_t launch_sinkhorn_fused32(const double* C, double* u, double* v, double* scratch, int B, int N, int M,
                                           int iters, cudaStream_t st) {
    const int R1 = N + 1, C1 = M + 1;
    const int ldk64 = (M + 1) & ~1, ldk = (M + 3) & ~3, ldv = (C1 + 1) & ~1;
    const int RS = (N + SKF_CLUSTER - 1) / SKF_CLUSTER;           // real rows per CTA (the dustbin row needs no storage)
    float* Kg = reinterpret_cast<float*>(scratch);
    int* flags = reinterpret_cast<int*>(scratch + (size_t)B * R1 * ldk64);
    cudaError_t e;
    cudaLaunchConfig_t cfg1, cfg2;
    cudaLaunchAttribute attr1[1], attr2[1];
    int rows1 = 0, rows2 = 0, cl1 = 0, cl2 = 0;
    static const int force = [] { const char* v = getenv("MDGAT_SK_CTAS"); return v ? atoi(v) : 0; }();
    // MDGAT_SK_THREADS: 512 (16 warps, 4 rows per warp round, one thread per column) or 1024 (32 warps, 2 rows, two threads per column)
    static const int threads = [] { const char* v = getenv("MDGAT_SK_THREADS"); return v ? atoi(v) : SK32_DEFAULT_THREADS; }();
    // Early exit once no b_j moved by more than tol relative (MDGAT_SK_TOL, 0 = "repeats bit for bit"): see the kernel comment
    static const double tol = [] { const char* v = getenv("MDGAT_SK_TOL"); return v ? atof(v) : SK32_EXIT_TOL; }();
    bool two = false;
    if (force == 2) {
        if (sk32_config<2, 512>(cfg2, attr2, B, RS, ldk, ldv, st, rows2, cl2) == cudaSuccess && cl2 > 0) two = true;
        else (void)cudaGetLastError();
    }
    const bool wide = !two && threads == 1024;
    if (!two && (e = wide ? sk32_config<1, 1024>(cfg1, attr1, B, RS, ldk, ldv, st, rows1, cl1)
                          : sk32_config<1, 512>(cfg1, attr1, B, RS, ldk, ldv, st, rows1, cl1)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(flags, 0, sizeof(int) * 2 * (size_t)B, st)) != cudaSuccess) return e;
    if (two) e = cudaLaunchKernelEx(&cfg2, sinkhorn_fused32_kernel<2, 512>, C, Kg, u, v, flags, N, M, iters, RS, rows2, ldk, tol);
    else if (wide) e = cudaLaunchKernelEx(&cfg1, sinkhorn_fused32_kernel<1, 1024>, C, Kg, u, v, flags, N, M, iters, RS, rows1, ldk, tol);
    else e = cudaLaunchKernelEx(&cfg1, sinkhorn_fused32_kernel<1, 512>, C, Kg, u, v, flags, N, M, iters, RS, rows1, ldk, tol);
    if (e != cudaSuccess) return e;
    sinkhorn_safe_kernel<<<B, 1024, 0, st>>>(C, u, v, flags, N, M, iters);
    count_launch(2);
    return cudaGetLastError();
}

cudaError_t launch_sinkhorn_fused(const double* C, double* u, double* v, double* scratch, int B, int N, int M,
                                  int iters, cudaStream_t st, bool k32) {
    if (k32) return launch_sinkhorn_fused32(C, u, v, scratch, B, N, M, iters, st);
    const int R1 = N + 1, C1 = M + 1;
    const int ldk = (M + 1) & ~1, ldv = (C1 + 1) & ~1;
    const int RS = (R1 + SKF_CLUSTER - 1) / SKF_CLUSTER;
    double* Kg = scratch;
    int* flags = reinterpret_cast<int*>(scratch + (size_t)B * R1 * ldk);
    int dev = 0, max_smem = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    const size_t fixed = ((size_t)3 * ldv + 3 * (size_t)((RS + 3) & ~3) + SKF_WARPS * SKF_NREG + 4) * sizeof(double);
    if (fixed + 1024 > (size_t)max_smem) return cudaErrorInvalidValue;
    int rows_smem = (int)(((size_t)max_smem - fixed) / ((size_t)ldk * sizeof(double)));
    if (rows_smem > RS) rows_smem = RS;
    const int nreg = (M <= SKF_THREADS) ? SKF_NREG : 0;
    const size_t smem = fixed + (size_t)rows_smem * ldk * sizeof(double);
    if ((e = cudaMemsetAsync(flags, 0, sizeof(int) * 2 * (size_t)B, st)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sinkhorn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(SKF_CLUSTER * B);
    cfg.blockDim = dim3(SKF_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = SKF_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if ((e = cudaLaunchKernelEx(&cfg, sinkhorn_fused_kernel, C, Kg, u, v, flags, N, M, iters, RS, rows_smem, nreg, ldk)) != cudaSuccess) return e;
    sinkhorn_safe_kernel<<<B, 1024, 0, st>>>(C, u, v, flags, N, M, iters);
    count_launch(2);
    return cudaGetLastError();
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
