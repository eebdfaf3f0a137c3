// Internal (C++) launcher interface between the kernel translation units and the C ABI
// in capi.cu. Not part of the public boundary (that is include/mdgat_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mdgat {

enum { EPI_PLAIN = 0, EPI_QKV = 1 };

// debug timeline buffer (device pointer, or null), see Tracer in common.cuh
extern long long* g_trace_dev;
extern int g_debug_flags;

// every kernel launch of this library is counted (bench.py reports it as gpu_launches)
void count_launch(int n = 1);

struct GemmParams {
    const double* A0; int lda0; int K0;     // first K0 input columns
    const double* A1; int lda1;             // remaining K - K0 columns (may be null)
    const double* W; int ldw;               // [Nout][K] row-major, or tile-major blob weight if w_tiled
    int w_tiled;
    const double* bias;                     // [Nout] or null
    const double* Res; int ldres;           // [R][Nout] or null (may alias Y)
    double* Y; int ldy;
    int R, Nout, K;
    double scale; int relu;
    long long sA, sW, sY;                   // batch strides (elements) for blockIdx.z
    // EPI_QKV: scatter the 384 output columns into head-major Q/K/V buffers
    double *Qh, *Kh, *Vh; int rows0, n0, n1;
};
cudaError_t launch_gemm(const GemmParams& p, int epi, int batch, cudaStream_t st);

// Per-side operands of one attention launch (side 0 and side 1 of a GNN layer share one grid).
struct AttnSides {
    const double* Q[2]; const double* K[2]; const double* V[2];   // head-major (B,4,n,36/36/34)
    double* Out[2];                                              // message rows (B*N) x ldo, or logits (B,4,N,M)
    int N[2], M[2];                                              // queries / sources of that side
};
// Full-softmax multi-head attention (flash-style, fp64 DMMA), nsides in {1, 2}.
cudaError_t launch_attention_full(const AttnSides& ps, int B, int nsides, int ldo, cudaStream_t st);
// Dense scaled logits S (B,4,N,M) = q.k / sqrt(32) (the QK^T half of the flash kernel, stored).
cudaError_t launch_attention_logits(const AttnSides& ps, int B, int nsides, cudaStream_t st);
// Exact top-k selection + softmax + sparse P.V from materialised logits S (B,4,N,M).
cudaError_t launch_topk_softmax_pv(const double* S, const double* V, double* Out, int ldo,
                                   int B, int N, int M, int topk, cudaStream_t st);
// dynamic_attention() in one persistent kernel (DMMA logits producers + selection / P.V consumers per CTA): ring = scratch of
// at least topk_fused_ring_doubles(B, N, M) doubles; topk_fused_supported() says whether the shape qualifies (else the
// two launches above / below run)
size_t topk_fused_ring_doubles(int B, int N, int M);
size_t topk_fused_ring_need(const AttnSides& ps, int B, int nsides);
bool topk_fused_supported(int nsides, const int* N, const int* M, int topk);
cudaError_t launch_topk_fused(const AttnSides& ps, int B, int nsides, int ldo, int topk, double* ring, size_t ring_doubles, cudaStream_t st);
// both sides of a layer in ONE launch (S[s]: dense logits (B,4,N[s],M[s]) of side s, V[s]: its head-major source values)
cudaError_t launch_topk_softmax_pv_sides(const double* const* S, const double* const* V, double* const* Out, int ldo,
                                         int B, const int* N, const int* M, int nsides, int topk, cudaStream_t st);

// Exact top-k threshold per logits row (B,4,N,M): kept set = { z > thr or (z == thr and column <= jlast) }, plus the row maximum
cudaError_t launch_topk_threshold(const double* S, double* thr, int* jlast, double* rmax, int B, int N, int M, int topk,
                                  cudaStream_t st);

// Encoder input staging: Xk (R x 4) = [x,y,z,score], Xd (R x 36) = [33 desc | 0 0 0]
cudaError_t launch_pack_inputs(const void* kpts0, const void* kpts1, const void* desc0, const void* desc1,
                               const void* sc0, const void* sc1, int in_dtype, int score_dtype,
                               int B, int N, int M, double* Xk, double* Xd, int* bad, cudaStream_t st);

// Sinkhorn (general, global-memory resident couplings)
cudaError_t launch_fill_dustbin(double* C, const double* bin_score, int B, int N, int M, cudaStream_t st);
cudaError_t launch_sinkhorn(const double* C, double* u, double* v, int B, int N, int M, int iters, cudaStream_t st);
// Fused: one launch, one 8-CTA cluster per pair, kernel matrix in distributed shared memory.
size_t sinkhorn_scratch_doubles(int B, int N, int M);
// k32: kernel matrix stored in float32 (sums and potentials in float64), see sinkhorn_fused32_kernel
cudaError_t launch_sinkhorn_fused(const double* C, double* u, double* v, double* scratch, int B, int N, int M,
                                  int iters, cudaStream_t st, bool k32 = false);

// Backward of attention() / dynamic_attention() (training path, attention_bwd.cu): head-major Qh, Kh (ld 36), Vh (ld 34), message
// O and its gradient dO as (B,4,N,32); outputs dQ (B,4,N,32), dK, dV (B,4,M,32)
size_t attention_bwd_scratch_doubles(int B, int N, int M, int topk);
cudaError_t launch_attention_backward(const double* Qh, const double* Kh, const double* Vh, const double* O, const double* dO,
                                      double* dQ, double* dK, double* dV, int B, int N, int M, int topk, double* scratch,
                                      cudaStream_t st);

// Backward of log_optimal_transport (training path, sinkhorn_bwd.cu): gC = dL/d(couplings) from G = dL/dZ; the int behind the
// scratch doubles is set when a row of the couplings spans more than 600 (scaling form not representable)
size_t sinkhorn_bwd_scratch_doubles(int B, int N, int M, int iters);
cudaError_t launch_sinkhorn_backward(const double* C, const double* G, double* gC, double* scratch, int B, int N, int M,
                                     int iters, cudaStream_t st);

struct MatchParams {
    const double* C; const double* u; const double* v;
    int B, N, M;
    int match_mode, mutual_check; double match_threshold;
    int loss_mode; double gamma;
    const int16_t* gt0; const int16_t* gt1;
    int64_t* matches0; int64_t* matches1; double* ms0; double* ms1;
    double* loss; int* nvalid0; double* Z;
    double* scratch;
    const int* bad;                 // per pair: an input was NaN / Inf (launch_pack_inputs), or nullptr
};
cudaError_t launch_match_extract(const MatchParams& p, cudaStream_t st);

cudaError_t launch_knn(const double* x, const double* src, int64_t* idx, int B, int n, int m, int k, cudaStream_t st);

// float64-faithful GEMM on tcgen05 int8 tensor cores (Ozaki splitting), see ozaki_gemm.cu
struct OzGemmArgs {
    const int8_t* Xs[2]; const double* rowscale[2];   // per 128-column k chunk, from launch_slice_rows
    const int8_t* Ws; const double* colscale;     // from packing.slice_weight
    const double* bias; const double* Res; int ldres;
    double* Y; int ldy;
    int R, Nout, K, relu, epi;
    double *Qh, *Kh, *Vh; int rows0, n0, n1;
    int8_t* slice_out; double* slice_scale;       // optional: digit planes + row scales of Y for the next GEMM (ozaki_gemm_can_slice)
};
size_t ozaki_slices_bytes(int R, int K, int S);
bool ozaki_gemm_can_slice(int R, int Nout);
// chunk c (128 columns) of the input goes to Xs + c * ozaki_slices_bytes(R, 128, S) and rowscale + c * Rpad
cudaError_t launch_slice_rows(const double* A0, int ld0, int K0, const double* A1, int ld1, int K1, int R, int S,
                              int8_t* Xs, double* rowscale, cudaStream_t st);
cudaError_t launch_ozaki_gemm(const OzGemmArgs& a, int S, cudaStream_t st);
cudaError_t measure_i8_peak(double* tops);

// float64-faithful attention on tcgen05 int8 tensor cores (digit planes of q/k/v, exact int32 accumulation in
// TMEM), see attention_i8.cu. One AttnI8Side holds the digit planes of ONE side's q, k and v head vectors.
struct AttnI8Side {
    int8_t *Qs, *Ks, *Vs;                 // Q planes per 128-row query tile, K / V^T planes per 32-row source tile
    double *qscale, *vscale;              // per query row, per (b, h, channel)
    int* kexp;                            // per source row: exponent of its scale, shifted into the float64 exponent field
    float *kscale_f, *ktilemax;           // fp32 copy of kscale, largest kscale per source tile
    int n;                                // keypoints of this side
    int S;                                // digit planes per operand (4..7)
};
size_t attn_i8_side_bytes(int B, int n, int S);
AttnI8Side attn_i8_carve(void* base, int B, int n, int S);
bool attn_i8_supported(int N, int M);
cudaError_t launch_attn_i8_slice(const double* Qh, const double* Kh, const double* Vh, const AttnI8Side& o, int B, cudaStream_t st);
// both sides of a layer (index 0 / 1) in one launch
cudaError_t launch_attn_i8_slice_sides(const double* const* Qh, const double* const* Kh, const double* const* Vh, const AttnI8Side* o,
                                       int B, cudaStream_t st);
// SP: byte planes of the probabilities; supported (S, SP): (4,3) (4,4) (5,4) (6,5) (7,6)
// mode: AI_MODE_FULL softmax over all sources; AI_MODE_LOGITS store the dense scaled logits (B,4,N,M) into Out;
// AI_MODE_TOPK softmax over the kept set described by tk (from launch_topk_threshold on the logits of AI_MODE_LOGITS)
enum { AI_MODE_FULL = 0, AI_MODE_LOGITS = 1, AI_MODE_TOPK = 2 };
struct AttnI8TopK { const double* thr[2]; const int* jlast[2]; const double* rmax[2]; };   // per side, rows in (B,4,N) order
// mp != nullptr (AI_MODE_FULL / AI_MODE_TOPK): the messages leave the kernel as the int8 digit planes the next GEMM
// reads (ozaki_gemm.cu operand layout, S planes, rows row0[s] + b * N[s] + i) instead of float64 rows; the digit scale
// of a row is the per-pair bound 2^e >= max |v| of its source values (a message is a convex combination of them),
// rowscale = 2^(e - 12). Out is not written then.
struct AttnI8MsgPlanes { int8_t* Xs; double* rowscale; int S; long long row0[2]; };
cudaError_t launch_attn_i8(const AttnI8Side* q, const AttnI8Side* kv, double* const* Out, int B, int nsides, int ldo,
                           int mode, const AttnI8TopK* tk, int SP, cudaStream_t st, const AttnI8MsgPlanes* mp = nullptr);

// Batched Kabsch registration + match statistics (one CTA per pair)
cudaError_t launch_register_pairs(const void* kpts0, const void* kpts1, int kp_dtype, const int64_t* matches0,
                                  const int16_t* gt0, const double* T_gt, int B, int N, int M,
                                  double* T_out, double* stats, cudaStream_t st);

// Ground-truth matches / T_gt / repeatability of a batch of pairs from poses and calibration
size_t prepare_pairs_smem(int N, int M);
cudaError_t launch_prepare_pairs(const double* kp1, const double* kp2, const double* pose1, const double* pose2,
                                 const double* calib, int calib_per_pair, int B, int N, int M, double threshold,
                                 int mutual_check, int16_t* match1, int16_t* match2, double* T_gt, int* rep,
                                 cudaStream_t st);

cudaError_t measure_fp64_peak(double* dmma_tflops, double* dfma_tflops);
cudaError_t measure_fp64_mixed(double* dmma_tflops, double* dfma_tflops);
cudaError_t measure_dmma_tiled(double* tflops);

}  // namespace mdgat
