/* mdgat_b200.h -- C ABI of the B200-native MDGAT-matcher hot path (libmdgat_b200.so).
 *
 * The reference exposes this path as a Python torch.nn.Module
 * (/root/reference/models/mdgat.py:315-603 MDGAT, models/superglue.py:315-625 SuperGlue);
 * it has no native code, so these entry points are what a ctypes/cffi binding inside the
 * reference's own models/mdgat.py would call in place of its eager PyTorch ops
 * (INTEGRATION.md shows that stub). Plain pointers and sizes only: every pointer named
 * d_* is a device pointer on the current CUDA device, every launch goes to `stream`
 * (a cudaStream_t passed as void*), nothing synchronises unless stated.
 *
 * All functions return MDGAT_OK (0) or a negative error; mdgat_last_error() has the text.
 * Activation buffers are point-major float64: one row per keypoint, 128 channels, row
 * stride MDGAT_LDX doubles; side 0 rows (B*N) come first, side 1 rows (B*M) after.
 */
#ifndef MDGAT_B200_H
#define MDGAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDGAT_OK 0
#define MDGAT_ERR_INVALID (-1)   /* bad argument (shape, alignment, k > M ...) */
#define MDGAT_ERR_CUDA (-2)      /* a CUDA runtime call failed */
#define MDGAT_ERR_WORKSPACE (-3) /* workspace too small */

#define MDGAT_LDX 132            /* row stride (doubles) of 128-channel activation rows */
#define MDGAT_DESC_IN 33         /* FPFH descriptor width (load_data.py:146-165) */

/* match-extraction variants (mdgat.py:442-483) */
#define MDGAT_MATCH_DUSTBIN 0    /* loss_method != 'superglue': arg-max including the dustbin */
#define MDGAT_MATCH_THRESHOLD 1  /* loss_method == 'superglue': inner arg-max + match_threshold */

/* loss variants computed on the device (mdgat.py:486-594) */
#define MDGAT_LOSS_NONE 0
#define MDGAT_LOSS_TRIPLET 1     /* mdgat.py:512-546 (test.py default): d_loss = 1 double, the batch mean */
#define MDGAT_LOSS_GAP 2         /* mdgat.py:547-594 (train.py default): d_loss = B doubles, one per pair; any N, M */
#define MDGAT_LOSS_SUPERGLUE 3   /* mdgat.py:487-511: d_loss = 1 double; gt with -1 = unmatched (NOT remapped); N == M */

/* GEMM engines for the per-layer projections (q/k/v, MLP) */
#define MDGAT_GEMM_DMMA_F64 0     /* mma.sync.m8n8k4.f64 (DMMA) on the FP64 pipe */
#define MDGAT_GEMM_TCGEN05_I8 1   /* float64-faithful Ozaki splitting on tcgen05.mma kind::i8 (TMEM accumulators) */

/* attention engines (full-softmax layers and the logits of the dynamic layers) */
#define MDGAT_ATTN_DMMA_F64 0     /* flash attention on DMMA (FP64 pipe) */
#define MDGAT_ATTN_TCGEN05_I8 1   /* float64-faithful digit products on tcgen05.mma kind::i8: Q K^T and P V in TMEM (full-attention
                                   * layers; the dense logits of the top-k layers come from the DMMA kernel, which is faster there) */
#define MDGAT_ATTN_TCGEN05_I8_ALL 2   /* the same, top-k layers included */

/* input element types */
#define MDGAT_F32 0
#define MDGAT_F64 1

const char* mdgat_last_error(void);
int mdgat_abi_version(void);

/* ---- packed weights ------------------------------------------------------------------
 * One float64 blob, BatchNorm folded into the preceding 1x1 conv, q/k/v output channels
 * and merge input channels permuted from the reference's interleaved c = d*4 + h
 * (mdgat.py:227) to head-major c' = h*32 + d. Every W[Cout][Cin] below is stored TILE-MAJOR:
 * [ceil(Cout/128)][ceil(Cin/32)][128][36] doubles, zero padded, so that one (128 x 32) GEMM stage is a
 * contiguous block for a single TMA bulk copy (packing.tile_weight). Order:
 *   kenc: W[32][4] b[32] W[64][32] b[64] W[128][64] b[128] W[128][128] b[128]
 *   denc: W[64][36] (33 inputs zero-padded to 36) b[64] W[128][64] b[128] W[128][128] b[128]
 *   per GNN layer (2L): Wqkv[384][128] bqkv[384] Wmlp0'[256][256] bmlp0'[256] Wmlp3[128][256] bmlp3[128]
 *                       (Wmlp0' = [W1x | W1m Wmerge], bmlp0' = b1 + W1m bmerge: the merge conv of
 *                        mdgat.py:237 composed with the first MLP conv of mdgat.py:248)
 *   final_proj: W[128][128] b[128];  bin_score (1 double, padded to 4)
 * mdgat_weight_blob_doubles(L) is the total length. Replaces MDGAT.__init__'s parameter
 * registration (mdgat.py:325-360) on the device side. */
size_t mdgat_weight_blob_doubles(int L);

/* ---- whole forward (mdgat.py:369-483 + triplet loss 512-546) ------------------------- */
typedef struct {
    int B, N, M;              /* batch, keypoints in set 0 / set 1 */
    int L;                    /* 2L GNN layers, names ['self','cross']*L (mdgat.py:353) */
    int sinkhorn_iters;       /* config['sinkhorn_iterations'] */
    const int* layer_k;       /* host array [2L]: top-k of layer i, 0 = full attention
                                 (schedule of mdgat.py:268-272 evaluated by the caller) */
    int match_mode;           /* MDGAT_MATCH_* */
    int mutual_check;         /* config['mutual_check'] */
    double match_threshold;   /* config['match_threshold'] */
    int loss_mode;            /* MDGAT_LOSS_* */
    double triplet_gamma;     /* config['triplet_loss_gamma'] */
    int in_dtype;             /* MDGAT_F32 / MDGAT_F64: element type of kpts/desc inputs */
    int score_dtype;          /* element type of scores0/1 (the reference does not cast them) */
    int write_Z;              /* also materialise Z (B,N+1,M+1) into d_Z (debug / other losses) */
    int gemm_mode;            /* MDGAT_GEMM_* */
    int gemm_slices;          /* int8 digit planes (7 bits each) per GEMM operand in MDGAT_GEMM_TCGEN05_I8 mode: 4..7.
                                 5 (35 bits) is the smallest count that passes the 131 k-row parity sweep against the
                                 unmodified reference (0 index flips, score error <= 7e-8); 7 = 49 bits ("exact") */
    int attn_mode;            /* MDGAT_ATTN_* (tcgen05 needs N, M <= 4096; larger sets use the DMMA kernel) */
    int attn_slices;          /* int8 digit planes (8 bits each) of q, k, v in the tcgen05 attention: 4..7 (0 = 7) */
    int attn_p_slices;        /* byte planes of the softmax probabilities: (attn_slices, attn_p_slices) must be one of
                                 (4,3) (4,4) (5,4) (6,5) (7,6); 0 = attn_slices - 1 */
    int sinkhorn_k32;         /* 1: the Sinkhorn kernel matrix exp(C_ij - max_j C_ij) is STORED in float32 (every sum,
                                 division and potential stays float64; potentials move by O(1e-7)) and the loop stops once
                                 no column scaling moved by more than 2^-35 relative in an iteration (skipped iterations
                                 bounded by (sinkhorn_iters - t) 2^-35 in the log-potentials; mdgat_forward_sinkhorn_status
                                 reports t per pair); 0: float64 storage, stop only when the iterate repeats bit for bit */
    int late_from;            /* > 0: GNN layers late_from .. 2L-1 use the late_* digit-plane counts (never more planes than
                                 the first setting). Experimental and off by default: 4/4/4 from layer 7 on keeps the match
                                 indices on the 131 k-row sweep but its worst score error is 6.9e-4. 0: one
                                 setting for all layers */
    int late_gemm_slices, late_attn_slices, late_attn_p_slices;
    const void* d_weights_i8_late;   /* device: the weight digit planes packed with late_gemm_slices (all layers), or NULL */
} mdgat_forward_cfg;

typedef struct {
    const void* d_kpts0;      /* (B,N,3) */
    const void* d_kpts1;      /* (B,M,3) */
    const void* d_desc0;      /* (B,N,33) */
    const void* d_desc1;      /* (B,M,33) */
    const void* d_scores0;    /* (B,N) */
    const void* d_scores1;    /* (B,M) */
    const int16_t* d_gt0;     /* (B,N) int16, "no match" = M (mdgat.py:519) or -1 (both accepted; SUPERGLUE needs -1), or NULL */
    const int16_t* d_gt1;     /* (B,M) int16, "no match" = N or -1, or NULL */
} mdgat_forward_in;

typedef struct {
    int64_t* d_matches0;      /* (B,N) int64, -1 = invalid */
    int64_t* d_matches1;      /* (B,M) */
    double* d_mscores0;       /* (B,N) */
    double* d_mscores1;       /* (B,M) */
    double* d_loss;           /* loss_mode TRIPLET / SUPERGLUE: 1 double; GAP: B doubles (one per pair); untouched for NONE */
    int* d_nvalid0;           /* 1 int: number of valid matches0 (the reference branches on it, mdgat.py:465) */
    double* d_Z;              /* (B,N+1,M+1) or NULL unless write_Z */
} mdgat_forward_out;

size_t mdgat_forward_workspace_bytes(const mdgat_forward_cfg* cfg);
/* d_weights_i8: int8-sliced copy of the per-layer GEMM weights (packing.pack_state_dict_i8), required in
 * MDGAT_GEMM_TCGEN05_I8 mode, NULL otherwise. */
int mdgat_forward(const mdgat_forward_cfg* cfg, const double* d_weights, const void* d_weights_i8,
                  const mdgat_forward_in* in, const mdgat_forward_out* out,
                  void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- individual operators (unit-parity surface; same kernels the forward uses) -------- */

/* Y[r][n] = act(scale * sum_k X[r][k] W[n][k] + bias[n]) + Res[r][n]
 * X is the concatenation [X0 (K0 cols) | X1 (K1 cols)]; X1/bias/Res may be NULL.
 * Replaces nn.Conv1d(k=1) + folded BatchNorm1d + ReLU of MLP() (mdgat.py:34-46). */
int mdgat_linear_f64(const double* d_X0, int ldx0, int K0, const double* d_X1, int ldx1, int K1,
                     const double* d_W, int ldw, const double* d_bias, const double* d_Res, int ldres,
                     double* d_Y, int ldy, int R, int Nout, double scale, int relu, void* stream);

/* Same contract as mdgat_linear_f64 (no scale), computed on the tcgen05 int8 tensor cores: X is split on
 * the device into `slices` int8 digit planes per row, W arrives pre-split (packing.slice_weight: planes in
 * UMMA canonical K-major order + colscale = 2^f_n); the int32 TMEM accumulators of the slice products are
 * recombined in float64. K0 (+K1) in {128, 256, 512}, Nout multiple of 64. */
size_t mdgat_linear_i8_scratch_bytes(int R, int K, int slices);
int mdgat_linear_i8(const double* d_X0, int ldx0, int K0, const double* d_X1, int ldx1, int K1,
                    const void* d_Wslices, const double* d_colscale, const double* d_bias, const double* d_Res, int ldres,
                    double* d_Y, int ldy, int R, int Nout, int relu, int slices, void* d_scratch, void* stream);

/* Batched Y[z] = scale * X[z] W[z]^T, z < batch (element strides sX, sW, sY).
 * Replaces torch.einsum('bdn,bdm->bnm') / sqrt(D) (mdgat.py:430-431) and the dense logits
 * einsum of dynamic_attention (mdgat.py:201). */
int mdgat_gemm_nt_f64(const double* d_X, int ldx, long long sX, const double* d_W, int ldw, long long sW,
                      double* d_Y, int ldy, long long sY, int R, int Nout, int K, int batch,
                      double scale, void* stream);

/* Encoders: denc(desc) + kenc(kpts, scores) for rows of both sides (mdgat.py:176-188,144-155,392-393).
 * d_X out: (B*N + B*M) x MDGAT_LDX. d_tmp: scratch of mdgat_encode_scratch_doubles(R). */
size_t mdgat_encode_scratch_doubles(int R);
int mdgat_encode(const mdgat_forward_in* in, int B, int N, int M, int in_dtype, int score_dtype,
                 const double* d_weights, double* d_X, double* d_tmp, void* stream);

/* Multi-head attention message for one side (mdgat.py:190-194 / 196-210), head-major inputs:
 * d_Q (B,4,N,36), d_K (B,4,M,36), d_V (B,4,M,34). topk == 0 -> softmax over all M,
 * else exactly-k selection (lowest index wins ties). d_logits: scratch of
 * mdgat_attention_f64_scratch_doubles(B, N, M) doubles (the dense (B,4,N,M) logits or the ring of the one-kernel
 * variant, whichever is larger), needed only when topk > 0. Output rows (B*N) x ldo, column h*32+d. */
size_t mdgat_attention_f64_scratch_doubles(int B, int N, int M);
int mdgat_attention_f64(const double* d_Q, const double* d_K, const double* d_V, double* d_Out, int ldo,
                        int B, int N, int M, int topk, double* d_logits, void* stream);

/* Same contract on the tcgen05 int8 tensor cores (attention_i8.cu): q, k, v are cut into base-256 digit planes on
 * the device, Q K^T and P V are exact int32 digit products in TMEM, recombined in float64.
 * d_scratch: mdgat_attention_i8_scratch_bytes(B, N, M) bytes. Requires N, M <= 4096. slices / p_slices: digit planes of
 * q, k, v / of the probabilities, as mdgat_forward_cfg.attn_slices / attn_p_slices. */
size_t mdgat_attention_i8_scratch_bytes(int B, int N, int M);   /* sized for 7 planes, enough for any setting */
int mdgat_attention_i8(const double* d_Q, const double* d_K, const double* d_V, double* d_Out, int ldo,
                       int B, int N, int M, int topk, double* d_logits, void* d_scratch, int slices, int p_slices,
                       void* stream);

/* log_optimal_transport (mdgat.py:279-308) on couplings already holding scores in [:N,:M]:
 * fills the dustbin row/column with bin_score (read from d_bin_score), runs `iters`
 * Sinkhorn iterations; leaves u (B,N+1), v (B,M+1) such that
 * Z = couplings + u + v - norm. fp64 throughout.
 * d_scratch != NULL (mdgat_sinkhorn_scratch_doubles): fused path, one launch, one 8-CTA
 * cluster per pair with the kernel matrix in distributed shared memory (what mdgat_forward
 * uses). d_scratch == NULL: one kernel per half-iteration reading the couplings from L2. */
size_t mdgat_sinkhorn_scratch_doubles(int B, int N, int M);
/* After a fused run (synchronises): per pair, whether the pair was redone by the log-domain
 * fallback (h_flags) and how many iterations ran before the fused kernel stopped (h_iters <= iters):
 * mdgat_sinkhorn_f64 stops once the iterate repeats bit for bit, every further iteration being a no-op;
 * mdgat_sinkhorn_f64_k32 stops once no column scaling moved by more than 2^-35 relative in an iteration
 * (the skipped iterations move a log-potential by at most (iters - h_iters) 2^-35; MDGAT_SK_TOL=0 in the
 * environment restores the bit-for-bit rule). */
int mdgat_sinkhorn_read_status(const double* d_scratch, int B, int N, int M, int* h_flags, int* h_iters);
/* The same status for the Sinkhorn stage of the last mdgat_forward() that ran on this workspace with this cfg
 * (the caller synchronises the forward's stream first). */
int mdgat_forward_sinkhorn_status(const mdgat_forward_cfg* cfg, const void* d_workspace, int* h_flags, int* h_iters);
int mdgat_sinkhorn_f64(double* d_couplings, const double* d_bin_score, double* d_u, double* d_v,
                       int B, int N, int M, int iters, double* d_scratch, void* stream);
/* Same contract (d_scratch required) with the kernel matrix stored in float32, float64 arithmetic (mdgat_forward_cfg.sinkhorn_k32). */
int mdgat_sinkhorn_f64_k32(double* d_couplings, const double* d_bin_score, double* d_u, double* d_v,
                           int B, int N, int M, int iters, double* d_scratch, void* stream);

/* Training path: hand-written backward of attention() / dynamic_attention() (mdgat.py:190-210), one side.
 * d_Q (B,4,N,36), d_K (B,4,M,36), d_V (B,4,M,34) head-major as for mdgat_attention_f64; d_O = the forward's message and
 * d_dO = its gradient as (B,4,N,32); outputs d_dQ (B,4,N,32), d_dK, d_dV (B,4,M,32). The probabilities are recomputed tile by
 * tile (nothing of size N x M is kept from the forward); for topk > 0 the kept set is rebuilt exactly (ties included) from
 * the dense logits of the forward's own kernel. d_scratch: mdgat_attention_backward_scratch_doubles(B, N, M, topk) doubles. */
size_t mdgat_attention_backward_scratch_doubles(int B, int N, int M, int topk);
int mdgat_attention_backward_f64(const double* d_Q, const double* d_K, const double* d_V, const double* d_O, const double* d_dO,
                                 double* d_dQ, double* d_dK, double* d_dV, int B, int N, int M, int topk, double* d_scratch,
                                 void* stream);

/* Training path: hand-written backward of log_optimal_transport (mdgat.py:279-308). d_couplings: the
 * (B,N+1,M+1) couplings WITH the dustbin row / column filled (what mdgat_sinkhorn_f64 leaves in place), d_gZ = dL/dZ of the
 * same shape, d_gcouplings (out) = dL/d(couplings) -- the caller takes [:, :N, :M] for the scores and the sum over the dustbin
 * row and column for bin_score. The iterates are recomputed in scaling form (2 T matrix-vector products), the reverse sweep is
 * 2 T more and one rank-2T contraction: O(T (N + M)) extra memory per pair where autograd retains 2 T (N+1)(M+1) tensors.
 * h_ill_conditioned (optional, host): set to 1 when a row of the couplings spans more than 600 and the result is invalid;
 * reading it synchronises the stream. */
size_t mdgat_sinkhorn_backward_scratch_doubles(int B, int N, int M, int iters);
int mdgat_sinkhorn_backward_f64(const double* d_couplings, const double* d_gZ, double* d_gcouplings, int B, int N, int M,
                                int iters, double* d_scratch, int* h_ill_conditioned, void* stream);

/* Match extraction + optional loss (MDGAT_LOSS_*) from (couplings, u, v) (mdgat.py:442-483, 487-594); Z is never formed. */
int mdgat_match_extract(const double* d_couplings, const double* d_u, const double* d_v,
                        int B, int N, int M, int match_mode, int mutual_check, double match_threshold,
                        int loss_mode, double gamma, const int16_t* d_gt0, const int16_t* d_gt1,
                        const mdgat_forward_out* out, double* d_scratch, void* stream);
size_t mdgat_match_scratch_doubles(int B, int N, int M);

/* knn() (mdgat.py:8-15): indices (B,n,k) int64 of the k nearest src points, nearest first.
 * x (B,3,n), src (B,3,m) float64 channel-major as in the reference. */
int mdgat_knn(const double* d_x, const double* d_src, int64_t* d_idx, int B, int n, int m, int k, void* stream);

/* Input side of the path: the per-item CPU work of SparseDataset.__getitem__
 * (load_data.py:213-292) batched on the device. kp1 (B,N,3) / kp2 (B,M,3) float64 in the LiDAR frame,
 * pose1/pose2 (B,4,4) cam0 poses, T_cam0_velo (4,4) shared or (B,4,4) per pair. Outputs: gt matches
 * (int16, -1 = none; nearest neighbour in world coordinates under `threshold`, optional mutual check),
 * T_gt (B,4,4) = inv(T_cam0_velo) inv(pose1) pose2 T_cam0_velo, rep[b] = number of set-1 points with a
 * neighbour closer than the threshold. */
int mdgat_prepare_pairs(const double* d_kp1, const double* d_kp2, const double* d_pose1, const double* d_pose2,
                        const double* d_T_cam0_velo, int calib_per_pair, int B, int N, int M, double threshold,
                        int mutual_check, int16_t* d_match1, int16_t* d_match2, double* d_T_gt, int* d_rep, void* stream);

/* Output side of the path: batched one-shot rigid registration from the
 * predicted matches and the match statistics the evaluation scripts accumulate, on the device.
 * Replaces solve_icp / calculate_error2 (utils/utils_test.py:73-110, 27-39) and the TP/FP/TN/FN
 * counting of test_registration_metric.py:216-246, which upstream run per pair on the CPU.
 * d_T: (B,4,4) transform taking the matched keypoints1 onto keypoints0 (R = U V^T, no reflection
 * fix, as upstream). d_stats: (B,8) = [n_valid, n_valid_gt, tp, fp, tn, fn, RTE, RRE]; RTE/RRE are
 * NaN without d_T_gt, the counts beyond n_valid are 0 without d_gt0 (gt value M or -1 = no match). */
int mdgat_register_pairs(const void* d_kpts0, const void* d_kpts1, int kp_dtype, const int64_t* d_matches0,
                         const int16_t* d_gt0, const double* d_T_gt, int B, int N, int M,
                         double* d_T, double* d_stats, void* stream);

/* Measured fp64 tensor-pipe peak of this device (DMMA.8x8x4 issue loop), TFLOP/s.
 * Synchronises. Used as the roofline denominator of the fp64 kernels. */
int mdgat_measure_fp64_peak(double* tflops_dmma, double* tflops_dfma);
/* Same, with DMMA and DFMA issued together (1 DMMA : 2 DFMA per warp): the rates each reaches in the mix.
 * tflops_dfma must have room for 2 doubles: [1] receives the DMMA rate of a register-tiled 4x4
 * outer-product loop at 16 warps/SM (the GEMM inner loop without its memory traffic). */
int mdgat_measure_fp64_mixed(double* tflops_dmma, double* tflops_dfma);
/* Issue-rate ceiling of tcgen05.mma kind::i8 on this device: one CTA per SM issues 128x256x32 MMAs on operands
 * resident in shared memory; result in int8 tera-operations (2 per multiply-add) per second. */
int mdgat_measure_i8_peak(double* tops);

/* ---- instrumentation (no reference counterpart; the reference has no tracing) ----
 * mdgat_launch_count: kernels launched by this library since load. Kernels replayed from a captured CUDA graph are not
 * launched through the library: the owner of the graph reports them with mdgat_launch_count_add (the number the counter
 * advanced by during the capture, once per replay).
 * Stage timers: when enabled, mdgat_forward brackets each run of same-stage launches with CUDA
 * events on the caller's stream; mdgat_profile_collect synchronises on the last event and returns
 * accumulated device milliseconds, launch counts and segment counts per stage, in the order
 * MDGAT_STAGE_* below. */
#define MDGAT_STAGE_ENCODE 0
#define MDGAT_STAGE_GEMM 1        /* q/k/v, merge, MLP, final_proj, score GEMMs */
#define MDGAT_STAGE_ATTN_FULL 2
#define MDGAT_STAGE_ATTN_TOPK 3   /* logits GEMM + selection/softmax/sparse PV */
#define MDGAT_STAGE_SINKHORN 4
#define MDGAT_STAGE_MATCH 5
#define MDGAT_STAGE_SLICE 6       /* digit-plane slicers of the tcgen05 engines */
#define MDGAT_STAGE_COUNT 7
long long mdgat_launch_count(void);
void mdgat_launch_count_add(long long n);
/* Debug timeline: d_buf = NULL (default, off) or a zeroed device buffer of 8 roles x (2 + 2*1024) int64. While set, CTA
 * (0,0,0) of the tcgen05 kernels records (tag, clock64) pairs per role (loader, MMA issuer, two epilogue warps):
 * buf[role*2050] = count, then the pairs. tools/trace_tcgen05.py prints the timeline. */
int mdgat_debug_trace(void* d_buf);
/* Debug switches for timeline experiments (0 = normal operation). Bit 0: the Ozaki GEMM epilogue skips its global
 * stores; bit 1: it skips the float64 recombination (results are wrong while 0 or 1 is set); bit 2: the epilogue warps
 * record timeline marks too (costs registers: their timing is then not that of the production kernel).
 * Bit 16 (0x10000, results unchanged): mdgat_forward lets the per-layer GEMMs cut their own outputs into digit planes
 * (same as MDGAT_FUSE_SLICE=1 in the environment). */
int mdgat_debug_flags(int flags);
int mdgat_profile_enable(int on);
int mdgat_profile_collect(double* ms, long long* launches, long long* segments, int n);

#ifdef __cplusplus
}
#endif
#endif /* MDGAT_B200_H */
