// C ABI of libmdgat_b200.so (declared in include/mdgat_b200.h) and the host-side
// orchestration of one forward pass: the layer loop of AttentionalGNN.forward
// (/root/reference/models/mdgat.py:259-276) and the op order of MDGAT.forward (:369-483),
// expressed as a fixed sequence of kernel launches on the caller's stream.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <vector>

#include "../../include/mdgat_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace mdgat_host {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace mdgat_host

using namespace mdgat;

// ---- launch counter and per-stage device timers (CUDA events on the caller's stream) ----
namespace {
// digit planes of the tcgen05 attention: (q/k/v planes, probability planes) from the cfg, 0 = defaults
inline int attn_planes(const mdgat_forward_cfg* cfg) { return cfg->attn_slices > 0 ? cfg->attn_slices : 7; }
inline int attn_p_planes(const mdgat_forward_cfg* cfg) { return cfg->attn_p_slices > 0 ? cfg->attn_p_slices : attn_planes(cfg) - 1; }
inline bool attn_planes_ok(int S, int SP) { return (S == 4 && (SP == 3 || SP == 4)) || (S >= 5 && S <= 7 && SP == S - 1); }
enum { ST_ENCODE = 0, ST_GEMM, ST_ATTN_FULL, ST_ATTN_TOPK, ST_SINKHORN, ST_MATCH, ST_SLICE, ST_COUNT };
struct Profiler {
    bool on = false;
    std::vector<cudaEvent_t> ev;       // ev[i] opens segment i; the last event closes the last segment
    std::vector<int> stage;
    std::vector<long long> launches_at;
    size_t used = 0;
    double ms[ST_COUNT] = {0};
    long long launches[ST_COUNT] = {0};
    long long segments[ST_COUNT] = {0};
};
Profiler g_prof;
std::atomic<long long> g_launches{0};

cudaEvent_t next_event() {
    if (g_prof.used == g_prof.ev.size()) { cudaEvent_t e; cudaEventCreate(&e); g_prof.ev.push_back(e); }
    return g_prof.ev[g_prof.used++];
}
// marks the beginning of a run of launches belonging to `stage` (ST_COUNT closes the forward)
void prof_mark(int stage, cudaStream_t st) {
    if (!g_prof.on) return;
    cudaEventRecord(next_event(), st);
    g_prof.stage.push_back(stage);
    g_prof.launches_at.push_back(g_launches.load());
}
}  // namespace
namespace mdgat { bool pdl_enabled() { static const bool on = [] { const char* e = getenv("MDGAT_PDL"); return !(e && e[0] == '0'); }(); return on; } }
namespace mdgat { void count_launch(int n) { g_launches.fetch_add(n); } long long* g_trace_dev = nullptr; int g_debug_flags = 0; }

namespace {

// ---- packed weight blob offsets (doubles); must match mdgat_matcher_b200/packing.py ----
struct EncOffsets { size_t w[4], b[4]; };
struct LayerOffsets { size_t wqkv, bqkv, w1, b1, w2, b2; };

constexpr size_t KENC_DIMS[5] = {4, 32, 64, 128, 128};
constexpr size_t DENC_DIMS[4] = {36, 64, 128, 128};
constexpr size_t tiled(size_t nout, size_t k) { return ((nout + 127) / 128) * ((k + 31) / 32) * 128 * 36; }
constexpr size_t LAYER_DOUBLES = tiled(384, 128) + 384 + tiled(256, 256) + 256 + tiled(128, 256) + 128;

struct BlobLayout {
    EncOffsets kenc, denc;
    size_t layers0, wf, bf, bin, total;
    explicit BlobLayout(int L) {
        size_t o = 0;
        for (int i = 0; i < 4; ++i) { kenc.w[i] = o; o += tiled(KENC_DIMS[i + 1], KENC_DIMS[i]); kenc.b[i] = o; o += KENC_DIMS[i + 1]; }
        for (int i = 0; i < 3; ++i) { denc.w[i] = o; o += tiled(DENC_DIMS[i + 1], DENC_DIMS[i]); denc.b[i] = o; o += DENC_DIMS[i + 1]; }
        layers0 = o; o += (size_t)(2 * L) * LAYER_DOUBLES;
        wf = o; o += tiled(128, 128); bf = o; o += 128;
        bin = o; o += 4;
        total = o;
    }
    LayerOffsets layer(int l) const {
        LayerOffsets r; size_t o = layers0 + (size_t)l * LAYER_DOUBLES;
        r.wqkv = o; o += tiled(384, 128); r.bqkv = o; o += 384;
        r.w1 = o; o += tiled(256, 256); r.b1 = o; o += 256;
        r.w2 = o; o += tiled(128, 256); r.b2 = o; o += 128;
        return r;
    }
};

constexpr int LDHID = 260;   // row stride of the 256-wide MLP hidden buffer

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Workspace {
    double *X, *Xk, *Xd, *Qh, *Kh, *Vh, *Msg, *Mg, *Hd, *MD, *S, *C, *u, *v, *mscratch, *skscratch;
    double *tkthr, *tkmax; int* tkjl;                   // tcgen05 top-k layers: per (b, h, row) threshold, maximum, last tied column
    double *rsX, *rsM, *rsH; int8_t *xsX, *xsM, *xsH;   // tcgen05 path: int8 slice planes + row scales of X, Msg, Hd (2 chunks)
    AttnI8Side ai[2];                                   // tcgen05 attention: digit planes of q/k/v of side 0 / side 1
    void* ai_base[2];                                   // their memory (re-carved for layers with fewer planes)
    int* bad;                                          // per pair: non-finite input seen
    size_t bytes, S_doubles;
};

// Carves the workspace; with base == nullptr only computes the size.
Workspace carve(char* base, int B, int N, int M, bool need_logits, int i8_slices = 0, int attn_i8 = 0) {
    Workspace w;
    const size_t R = (size_t)B * N + (size_t)B * M;
    size_t off = 0;
    auto take = [&](size_t doubles) {
        double* p = base ? reinterpret_cast<double*>(base + off) : nullptr;
        off += align_up(doubles * sizeof(double), 256);
        return p;
    };
    w.X = take(R * LDX);
    w.Xk = take(R * 4);
    w.Xd = take(R * 36);
    w.Qh = take(R * HEADS * LDH_QK);
    w.Kh = take(R * HEADS * LDH_QK);
    w.Vh = take(R * HEADS * LDH_V);
    w.Msg = take(R * LDX);
    w.Mg = take(R * LDX);
    w.Hd = take(R * LDHID);
    w.MD = take(R * LDX);
    // logits of both sides at once: self layers need N*N + M*M, cross layers 2*N*M (never more); the fused top-k kernel
    // uses the same buffer as its ring
    {
        const size_t dense = (size_t)B * HEADS * ((size_t)N * N + (size_t)M * M), ringd = topk_fused_ring_doubles(B, N, M);
        w.S_doubles = need_logits ? (dense > ringd ? dense : ringd) : 0;
        w.S = take(w.S_doubles);
    }
    w.C = take((size_t)B * (N + 1) * (M + 1));
    w.u = take((size_t)B * (N + 1));
    w.v = take((size_t)B * (M + 1));
    w.mscratch = take(4 * R + 8);
    w.bad = reinterpret_cast<int*>(take((size_t)(B + 1) / 2 + 1));
    w.skscratch = take(sinkhorn_scratch_doubles(B, N, M));
    w.tkthr = take(need_logits && attn_i8 ? R * HEADS : 0);
    w.tkmax = take(need_logits && attn_i8 ? R * HEADS : 0);
    w.tkjl = reinterpret_cast<int*>(take(need_logits && attn_i8 ? (R * HEADS + 1) / 2 : 0));
    const size_t Rpad = (R + 127) / 128 * 128;
    w.rsX = take(i8_slices ? Rpad : 0);
    w.rsM = take(i8_slices ? Rpad : 0);
    w.rsH = take(i8_slices ? 2 * Rpad : 0);
    const size_t chunk8 = i8_slices ? (ozaki_slices_bytes((int)R, 128, i8_slices) + 7) / 8 : 0;
    w.xsX = reinterpret_cast<int8_t*>(take(chunk8));
    w.xsM = reinterpret_cast<int8_t*>(take(chunk8));
    w.xsH = reinterpret_cast<int8_t*>(take(2 * chunk8));
    for (int s = 0; s < 2; ++s) {
        const int n = s == 0 ? N : M;
        void* p = take(attn_i8 ? attn_i8_side_bytes(B, n, attn_i8) / 8 : 0);
        w.ai_base[s] = p;
        if (attn_i8 && base) w.ai[s] = attn_i8_carve(p, B, n, attn_i8);
    }
    w.bytes = off;
    return w;
}

// w_tiled: W is a tile-major blob weight (TMA bulk staging); otherwise row-major [Nout][ldw]
cudaError_t linear(const double* X0, int ld0, int K0, const double* X1, int ld1, int K1,
                   const double* W, int ldw, const double* bias, const double* Res, int ldres,
                   double* Y, int ldy, int R, int Nout, double scale, int relu, cudaStream_t st, int w_tiled = 1) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.w_tiled = w_tiled;
    p.A0 = X0; p.lda0 = ld0; p.K0 = K0; p.A1 = X1; p.lda1 = ld1;
    p.W = W; p.ldw = ldw; p.bias = bias; p.Res = Res; p.ldres = ldres; p.Y = Y; p.ldy = ldy;
    p.R = R; p.Nout = Nout; p.K = K0 + K1; p.scale = scale; p.relu = relu;
    return launch_gemm(p, EPI_PLAIN, 1, st);
}

cudaError_t gemm_nt(const double* X, int ldx, long long sX, const double* W, int ldw, long long sW,
                    double* Y, int ldy, long long sY, int R, int Nout, int K, int batch, double scale,
                    cudaStream_t st) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.A0 = X; p.lda0 = ldx; p.K0 = K; p.W = W; p.ldw = ldw; p.Y = Y; p.ldy = ldy;
    p.R = R; p.Nout = Nout; p.K = K; p.scale = scale; p.sA = sX; p.sW = sW; p.sY = sY;
    return launch_gemm(p, EPI_PLAIN, batch, st);
}

// Messages of one GNN layer. nsides = 2: side 0 and side 1 in the same launches.
// qd / kvd != nullptr: tcgen05 engine; the digit planes of the query / source side of each grid side are already cut.
// tk: scratch of the tcgen05 top-k path (threshold / maximum / last tied column per (b, h, row) of every side)
struct TopKScratch { double* thr; double* rmax; int* jlast; };
cudaError_t attention_layer(const AttnSides& ps, int B, int nsides, int ldo, int topk, double* S, cudaStream_t st,
                            const AttnI8Side* qd = nullptr, const AttnI8Side* kvd = nullptr, int SP = 0,
                            const TopKScratch* tks = nullptr, const AttnI8MsgPlanes* mp = nullptr, bool* planes_written = nullptr,
                            size_t S_doubles = 0) {
    if (planes_written) *planes_written = false;
    if (qd) {
        if (topk <= 0) {
            if (planes_written) *planes_written = mp != nullptr;
            return launch_attn_i8(qd, kvd, ps.Out, B, nsides, ldo, AI_MODE_FULL, nullptr, SP, st, mp);
        }
        // dynamic_attention() (mdgat.py:196-210) on the tensor cores: dense logits (digit-plane Q K^T, stored), the exact
        // top-k threshold of every row, then the same kernel again as a masked softmax . V that recomputes the logits bit
        // for bit and zeroes the probabilities outside the kept set
        double* lg[2]; double* sp = S;
        for (int s = 0; s < nsides; ++s) { lg[s] = sp; sp += (size_t)B * HEADS * ps.N[s] * ps.M[s]; }
        cudaError_t e = launch_attn_i8(qd, kvd, lg, B, nsides, 0, AI_MODE_LOGITS, nullptr, SP, st);
        if (e != cudaSuccess) return e;
        if (tks == nullptr) {
            return launch_topk_softmax_pv_sides(lg, ps.V, ps.Out, ldo, B, ps.N, ps.M, nsides, topk, st);
        }
        AttnI8TopK tk;
        size_t off = 0;
        for (int s = 0; s < nsides; ++s) {
            tk.thr[s] = tks->thr + off; tk.rmax[s] = tks->rmax + off; tk.jlast[s] = tks->jlast + off;
            e = launch_topk_threshold(lg[s], tks->thr + off, tks->jlast + off, tks->rmax + off, B, ps.N[s], ps.M[s], topk, st);
            if (e != cudaSuccess) return e;
            off += (size_t)B * HEADS * ps.N[s];
        }
        if (planes_written) *planes_written = mp != nullptr;
        return launch_attn_i8(qd, kvd, ps.Out, B, nsides, ldo, AI_MODE_TOPK, &tk, SP, st, mp);
    }
    if (topk <= 0) return launch_attention_full(ps, B, nsides, ldo, st);
    // MDGAT_TOPK_FUSED=1 (or debug flag bit 17, per call, for the tests): one persistent kernel whose logits go through a
    // per-CTA ring instead of the dense scratch. Parity-green but measured slower than the two launches below
    static const bool fused_env = [] { const char* e = getenv("MDGAT_TOPK_FUSED"); return e && e[0] == '1'; }();
    if ((fused_env || (g_debug_flags & 0x20000) != 0) && topk_fused_supported(nsides, ps.N, ps.M, topk) && S_doubles >= topk_fused_ring_need(ps, B, nsides))
        return launch_topk_fused(ps, B, nsides, ldo, topk, S, S_doubles, st);
    // dense logits q.k / sqrt(32) for every (b, h) (mdgat.py:201), then exact-k selection per row
    AttnSides lg = ps;
    double* sp = S;
    for (int s = 0; s < nsides; ++s) { lg.Out[s] = sp; sp += (size_t)B * HEADS * ps.N[s] * ps.M[s]; }
    cudaError_t e = launch_attention_logits(lg, B, nsides, st);
    if (e != cudaSuccess) return e;
    return launch_topk_softmax_pv_sides(lg.Out, ps.V, ps.Out, ldo, B, ps.N, ps.M, nsides, topk, st);
}

cudaError_t encode(const mdgat_forward_in* in, int B, int N, int M, int in_dtype, int score_dtype,
                   const double* Wt, const BlobLayout& lay, double* X, double* Xk, double* Xd,
                   double* T0, double* T1, double* T2, int ldt2, cudaStream_t st, int* bad = nullptr) {
    const int R = B * N + B * M;
    cudaError_t e = launch_pack_inputs(in->d_kpts0, in->d_kpts1, in->d_desc0, in->d_desc1, in->d_scores0,
                                       in->d_scores1, in_dtype, score_dtype, B, N, M, Xk, Xd, bad, st);
    if (e != cudaSuccess) return e;
    // KeypointEncoder: 4 -> 32 -> 64 -> 128 -> 128 (mdgat.py:181), BN folded, ReLU on the first three
    if ((e = linear(Xk, 4, 4, nullptr, 0, 0, Wt + lay.kenc.w[0], 4, Wt + lay.kenc.b[0], nullptr, 0, T0, LDX, R, 32, 1.0, 1, st))) return e;
    if ((e = linear(T0, LDX, 32, nullptr, 0, 0, Wt + lay.kenc.w[1], 32, Wt + lay.kenc.b[1], nullptr, 0, T1, LDX, R, 64, 1.0, 1, st))) return e;
    if ((e = linear(T1, LDX, 64, nullptr, 0, 0, Wt + lay.kenc.w[2], 64, Wt + lay.kenc.b[2], nullptr, 0, T0, LDX, R, 128, 1.0, 1, st))) return e;
    if ((e = linear(T0, LDX, 128, nullptr, 0, 0, Wt + lay.kenc.w[3], 128, Wt + lay.kenc.b[3], nullptr, 0, T1, LDX, R, 128, 1.0, 0, st))) return e;
    // DescriptorEncoder: 33(36) -> 64 -> 128 -> 128 (mdgat.py:148); the last layer adds kenc (mdgat.py:392)
    if ((e = linear(Xd, 36, 36, nullptr, 0, 0, Wt + lay.denc.w[0], 36, Wt + lay.denc.b[0], nullptr, 0, T0, LDX, R, 64, 1.0, 1, st))) return e;
    if ((e = linear(T0, LDX, 64, nullptr, 0, 0, Wt + lay.denc.w[1], 64, Wt + lay.denc.b[1], nullptr, 0, T2, ldt2, R, 128, 1.0, 1, st))) return e;
    return linear(T2, ldt2, 128, nullptr, 0, 0, Wt + lay.denc.w[2], 128, Wt + lay.denc.b[2], T1, LDX, X, LDX, R, 128, 1.0, 0, st);
}

}  // namespace

extern "C" {

const char* mdgat_last_error(void) { return mdgat_host::g_err; }
int mdgat_abi_version(void) { return 4; }

size_t mdgat_weight_blob_doubles(int L) { return BlobLayout(L).total; }

size_t mdgat_forward_workspace_bytes(const mdgat_forward_cfg* cfg) {
    bool need = false;
    for (int i = 0; i < 2 * cfg->L; ++i) need = need || (cfg->layer_k && cfg->layer_k[i] > 0);
    const bool ai = (cfg->attn_mode == MDGAT_ATTN_TCGEN05_I8 || cfg->attn_mode == MDGAT_ATTN_TCGEN05_I8_ALL) && attn_i8_supported(cfg->N, cfg->M) && attn_i8_supported(cfg->M, cfg->N);
    return carve(nullptr, cfg->B, cfg->N, cfg->M, need, cfg->gemm_mode == MDGAT_GEMM_TCGEN05_I8 ? cfg->gemm_slices : 0, ai ? attn_planes(cfg) : 0).bytes;
}

int mdgat_forward(const mdgat_forward_cfg* cfg, const double* d_weights, const void* d_weights_i8,
                  const mdgat_forward_in* in, const mdgat_forward_out* out, void* d_workspace, size_t workspace_bytes,
                  void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int B = cfg->B, N = cfg->N, M = cfg->M, L = cfg->L;
    MDGAT_REQUIRE(B > 0 && N > 0 && M > 0 && L > 0, "mdgat_forward: B, N, M, L must be positive (got %d %d %d %d)", B, N, M, L);
    MDGAT_REQUIRE(cfg->layer_k != nullptr, "mdgat_forward: layer_k is NULL");
    MDGAT_REQUIRE((long long)B * (N + M) < (1ll << 31) / LDHID, "mdgat_forward: batch too large for 32-bit row indexing");
    bool need = false;
    for (int i = 0; i < 2 * L; ++i) {
        const int k = cfg->layer_k[i];
        if (k > 0) {
            need = true;
            // side 0 attends over N (self) or M (cross) sources, side 1 over M or N: topk() needs k <= both
            MDGAT_REQUIRE(k <= N && k <= M, "selected index k out of range (layer %d: k=%d, N=%d, M=%d)", i, k, N, M);
            MDGAT_REQUIRE(N <= 2048 && M <= 2048, "top-k attention supports at most 2048 source keypoints (N=%d, M=%d)", N, M);
        }
    }
    MDGAT_REQUIRE(cfg->loss_mode >= MDGAT_LOSS_NONE && cfg->loss_mode <= MDGAT_LOSS_SUPERGLUE, "mdgat_forward: unknown loss_mode %d", cfg->loss_mode);
    if (cfg->loss_mode == MDGAT_LOSS_GAP) {
        MDGAT_REQUIRE(in->d_gt0 && in->d_gt1, "gap loss needs gt_matches0/1");
        MDGAT_REQUIRE(cfg->match_mode == MDGAT_MATCH_DUSTBIN, "gap loss is defined on the dustbin match variant");
    }
    if (cfg->loss_mode == MDGAT_LOSS_SUPERGLUE) {
        MDGAT_REQUIRE(in->d_gt0 && in->d_gt1, "superglue loss needs gt_matches0/1");
        MDGAT_REQUIRE(N == M, "superglue loss needs N == M (the reference indexes an N-shaped mask with M, mdgat.py:501)");
    }
    if (cfg->loss_mode == MDGAT_LOSS_TRIPLET) {
        MDGAT_REQUIRE(in->d_gt0 && in->d_gt1, "triplet loss needs gt_matches0/1");
        MDGAT_REQUIRE(N == M, "triplet loss needs N == M (the reference raises IndexError otherwise, mdgat.py:537)");
        MDGAT_REQUIRE(cfg->match_mode == MDGAT_MATCH_DUSTBIN, "triplet loss is defined on the dustbin match variant");
    }
    const bool i8 = cfg->gemm_mode == MDGAT_GEMM_TCGEN05_I8;
    const int S8 = cfg->gemm_slices;
    MDGAT_REQUIRE(!i8 || (d_weights_i8 != nullptr && S8 >= 4 && S8 <= 7), "tcgen05 int8 GEMM mode needs the sliced weight blob and 4..7 slices (got %d)", S8);
    const bool ai8_any = (cfg->attn_mode == MDGAT_ATTN_TCGEN05_I8 || cfg->attn_mode == MDGAT_ATTN_TCGEN05_I8_ALL) && attn_i8_supported(N, M) && attn_i8_supported(M, N);
    const int AS = attn_planes(cfg), ASP = attn_p_planes(cfg);
    MDGAT_REQUIRE(!ai8_any || attn_planes_ok(AS, ASP), "tcgen05 attention: unsupported digit planes (attn_slices %d, attn_p_slices %d)", AS, ASP);
    // second digit-plane setting from layer late_from on (0: none): never more planes than the first (the workspace is
    // sized for the first), its own weight blob
    const int late_from = (cfg->late_from > 0 && cfg->late_from < 2 * L) ? cfg->late_from : 0;
    const int S8L = late_from ? cfg->late_gemm_slices : S8;
    const int ASL = late_from && cfg->late_attn_slices > 0 ? cfg->late_attn_slices : AS;
    const int ASPL = late_from && cfg->late_attn_slices > 0 ? (cfg->late_attn_p_slices > 0 ? cfg->late_attn_p_slices : ASL - 1) : ASP;
    if (late_from) {
        MDGAT_REQUIRE(!i8 || (cfg->d_weights_i8_late != nullptr && S8L >= 4 && S8L <= S8), "late layers: need d_weights_i8_late and 4 <= late_gemm_slices <= gemm_slices (got %d)", S8L);
        MDGAT_REQUIRE(!ai8_any || (attn_planes_ok(ASL, ASPL) && ASL <= AS), "late layers: unsupported attention digit planes (%d, %d)", ASL, ASPL);
    }
    Workspace w = carve(reinterpret_cast<char*>(d_workspace), B, N, M, need, i8 ? S8 : 0, ai8_any ? AS : 0);
    if (w.bytes > workspace_bytes) {
        mdgat_host::set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
        return MDGAT_ERR_WORKSPACE;
    }
    const BlobLayout lay(L);
    const double* Wt = d_weights;
    const int R0 = B * N, R1 = B * M, R = R0 + R1;

    prof_mark(ST_ENCODE, st);
    MDGAT_CUDA_OK(encode(in, B, N, M, cfg->in_dtype, cfg->score_dtype, Wt, lay, w.X, w.Xk, w.Xd, w.Msg, w.Mg, w.Hd, LDHID, st, w.bad));

    // head-major buffers: side 0 occupies the first R0*4 rows, side 1 the rest
    const double* Q0 = w.Qh; const double* Q1 = w.Qh + (size_t)R0 * HEADS * LDH_QK;
    const double* K0 = w.Kh; const double* K1 = w.Kh + (size_t)R0 * HEADS * LDH_QK;
    const double* V0 = w.Vh; const double* V1 = w.Vh + (size_t)R0 * HEADS * LDH_V;

    // MDGAT_FUSE_SLICE=1: the per-layer GEMMs cut their own outputs into digit planes when one CTA owns whole rows
    // (R >= 148 row tiles). Bit-identical results; measured slower on B200 (slice stage -1.04 ms, GEMM stage +1.14 ms:
    // the FP64-bound slicing tail runs while the CTA's tensor pipe idles), so the stand-alone launches stay the default.
    static const bool fuse_env = [] { const char* e = getenv("MDGAT_FUSE_SLICE"); return e && e[0] == '1'; }();
    const bool fuse_slice = i8 && !late_from && (fuse_env || (g_debug_flags & 0x10000) != 0) && ozaki_gemm_can_slice(R, DMODEL);   // bit 16: same switch, per call (tests)
    AttnI8Side ai_late[2] = {w.ai[0], w.ai[1]};
    if (late_from && ai8_any && ASL != AS) {
        ai_late[0] = attn_i8_carve(w.ai_base[0], B, N, ASL);
        ai_late[1] = attn_i8_carve(w.ai_base[1], B, M, ASL);
    }
    const int S8_first = S8, ASP_first = ASP;
    const void* const d_weights_i8_first = d_weights_i8;
    for (int l = 0; l < 2 * L; ++l) {
        const bool late = late_from && l >= late_from;
        const int S8 = late ? S8L : S8_first, ASP = late ? ASPL : ASP_first;          // this layer's digit planes
        const AttnI8Side* ai = late ? ai_late : w.ai;
        const void* d_weights_i8 = late ? cfg->d_weights_i8_late : d_weights_i8_first;
        const LayerOffsets lo = lay.layer(l);
        const bool cross = (l & 1) != 0;                  // names = ['self','cross']*L (mdgat.py:353)
        const int k = cfg->layer_k[l];
        // top-k layers only need the dense logits: the DMMA flash pipeline produces them faster than the digit-plane
        // route (no slicing pass, no Horner per logit) unless the caller asks for tcgen05 everywhere
        const bool ai8 = ai8_any && (k <= 0 || cfg->attn_mode == MDGAT_ATTN_TCGEN05_I8_ALL);
        // q/k/v of both sides with the shared layer weights (mdgat.py:227-232, :270)
        // int8 blob of layer l: [qkv: 12 col tiles x 1 k chunk][mlp0: 8 x 2][mlp3: 4 x 2] slice tiles, then the column
        // scales qkv[384] mlp0[2][256] mlp3[2][128]
        const size_t slice_tile = (size_t)S8 * 32 * 128;                // bytes of one (column tile, k chunk) of W slices
        const unsigned char* Li8 = i8 ? reinterpret_cast<const unsigned char*>(d_weights_i8) + (size_t)l * (slice_tile * 36 + 1152 * 8) : nullptr;
        const double* cs8 = i8 ? reinterpret_cast<const double*>(Li8 + slice_tile * 36) : nullptr;
        prof_mark(i8 && !(fuse_slice && l > 0) ? ST_SLICE : ST_GEMM, st);
        if (i8) {
            // digit planes of the residual stream: cut by the previous layer's last GEMM when it can (fuse_slice)
            if (!(fuse_slice && l > 0)) MDGAT_CUDA_OK(launch_slice_rows(w.X, LDX, DMODEL, nullptr, 0, 0, R, S8, w.xsX, w.rsX, st));
            prof_mark(ST_GEMM, st);
            OzGemmArgs a;
            memset(&a, 0, sizeof(a));
            a.Xs[0] = w.xsX; a.rowscale[0] = w.rsX; a.Ws = reinterpret_cast<const int8_t*>(Li8); a.colscale = cs8;
            a.bias = Wt + lo.bqkv; a.R = R; a.Nout = 3 * DMODEL; a.K = DMODEL; a.epi = EPI_QKV;
            a.Qh = w.Qh; a.Kh = w.Kh; a.Vh = w.Vh; a.rows0 = R0; a.n0 = N; a.n1 = M;
            MDGAT_CUDA_OK(launch_ozaki_gemm(a, S8, st));
        } else {
            GemmParams p;
            memset(&p, 0, sizeof(p));
            p.A0 = w.X; p.lda0 = LDX; p.K0 = DMODEL; p.W = Wt + lo.wqkv; p.ldw = DMODEL; p.bias = Wt + lo.bqkv;
            p.R = R; p.Nout = 3 * DMODEL; p.K = DMODEL; p.scale = 1.0; p.w_tiled = 1;
            p.Qh = w.Qh; p.Kh = w.Kh; p.Vh = w.Vh; p.rows0 = R0; p.n0 = N; p.n1 = M;
            MDGAT_CUDA_OK(launch_gemm(p, EPI_QKV, 1, st));
        }
        prof_mark(ai8 ? ST_SLICE : (k > 0 ? ST_ATTN_TOPK : ST_ATTN_FULL), st);
        bool msg_planes = false;                          // the attention kernel wrote the message digit planes itself
        // messages: side 0 reads side (cross ? 1 : 0), side 1 the other way round (mdgat.py:263-266)
        AttnSides ps;
        ps.Q[0] = Q0; ps.K[0] = cross ? K1 : K0; ps.V[0] = cross ? V1 : V0; ps.Out[0] = w.Msg; ps.N[0] = N; ps.M[0] = cross ? M : N;
        ps.Q[1] = Q1; ps.K[1] = cross ? K0 : K1; ps.V[1] = cross ? V0 : V1; ps.Out[1] = w.Msg + (size_t)R0 * LDX; ps.N[1] = M; ps.M[1] = cross ? N : M;
        if (ai8) {
            // digit planes of this layer's q, k, v (both sides), then the tcgen05 kernel: side s reads the k / v planes
            // of side s (self) or 1 - s (cross)
            const double* const qs[2] = {Q0, Q1}; const double* const ks[2] = {K0, K1}; const double* const vs[2] = {V0, V1};
            if (N > 0 && M > 0) { MDGAT_CUDA_OK(launch_attn_i8_slice_sides(qs, ks, vs, ai, B, st)); }
            else {
                MDGAT_CUDA_OK(launch_attn_i8_slice(Q0, K0, V0, ai[0], B, st));
                MDGAT_CUDA_OK(launch_attn_i8_slice(Q1, K1, V1, ai[1], B, st));
            }
            prof_mark(k > 0 ? ST_ATTN_TOPK : ST_ATTN_FULL, st);
            const AttnI8Side qd[2] = {ai[0], ai[1]};
            const AttnI8Side kvd[2] = {cross ? ai[1] : ai[0], cross ? ai[0] : ai[1]};
            const TopKScratch tks = {w.tkthr, w.tkmax, w.tkjl};
            // with the tcgen05 GEMM engine the messages leave the attention kernel as the digit planes of the MLP's second
            // k chunk (MDGAT_MSG_PLANES=0: float64 rows + the stand-alone slicer, as for the DMMA attention engine)
            static const bool msg_planes_env = [] { const char* e = getenv("MDGAT_MSG_PLANES"); return !(e && e[0] == '0'); }();
            const AttnI8MsgPlanes mpl = {w.xsM, w.rsM, S8, {0, (long long)R0}};
            MDGAT_CUDA_OK(attention_layer(ps, B, 2, LDX, k, w.S, st, qd, kvd, ASP, &tks, (i8 && msg_planes_env) ? &mpl : nullptr, &msg_planes));
        } else {
            MDGAT_CUDA_OK(attention_layer(ps, B, 2, LDX, k, w.S, st, nullptr, nullptr, 0, nullptr, nullptr, nullptr, w.S_doubles));
        }
        prof_mark(i8 ? ST_SLICE : ST_GEMM, st);
        // the merge conv (mdgat.py:237) is folded into the first MLP conv by the weight packer
        // mlp(cat[x, message]) : 256 -> 256 (BN folded, ReLU) -> 128, then the residual (mdgat.py:248, :274)
        if (i8) {
            OzGemmArgs a;
            memset(&a, 0, sizeof(a));
            // k chunk 0 = x (its slice planes are the ones the q/k/v projection used), k chunk 1 = message
            if (!msg_planes) MDGAT_CUDA_OK(launch_slice_rows(w.Msg, LDX, DMODEL, nullptr, 0, 0, R, S8, w.xsM, w.rsM, st));
            prof_mark(ST_GEMM, st);
            a.Xs[0] = w.xsX; a.rowscale[0] = w.rsX; a.Xs[1] = w.xsM; a.rowscale[1] = w.rsM; a.Ws = reinterpret_cast<const int8_t*>(Li8 + slice_tile * 12); a.colscale = cs8 + 384;
            a.bias = Wt + lo.b1; a.Y = w.Hd; a.ldy = LDHID; a.R = R; a.Nout = 2 * DMODEL; a.K = 2 * DMODEL; a.relu = 1; a.epi = EPI_PLAIN;
            if (fuse_slice) { a.slice_out = w.xsH; a.slice_scale = w.rsH; }       // the GEMM cuts its own output for the next one
            MDGAT_CUDA_OK(launch_ozaki_gemm(a, S8, st));
            a.slice_out = nullptr; a.slice_scale = nullptr;
            if (!fuse_slice) {
                prof_mark(ST_SLICE, st);
                MDGAT_CUDA_OK(launch_slice_rows(w.Hd, LDHID, 2 * DMODEL, nullptr, 0, 0, R, S8, w.xsH, w.rsH, st));
                prof_mark(ST_GEMM, st);
            }
            a.Xs[0] = w.xsH; a.rowscale[0] = w.rsH; a.Xs[1] = w.xsH + ozaki_slices_bytes(R, 128, S8); a.rowscale[1] = w.rsH + (size_t)((R + 127) / 128 * 128);
            a.Ws = reinterpret_cast<const int8_t*>(Li8 + slice_tile * 28); a.colscale = cs8 + 896;
            a.bias = Wt + lo.b2; a.Res = w.X; a.ldres = LDX; a.Y = w.X; a.ldy = LDX; a.Nout = DMODEL; a.relu = 0;
            if (fuse_slice && l + 1 < 2 * L) { a.slice_out = w.xsX; a.slice_scale = w.rsX; }   // planes of x for the next layer
            MDGAT_CUDA_OK(launch_ozaki_gemm(a, S8, st));
        } else {
            MDGAT_CUDA_OK(linear(w.X, LDX, DMODEL, w.Msg, LDX, DMODEL, Wt + lo.w1, 2 * DMODEL, Wt + lo.b1, nullptr, 0, w.Hd, LDHID, R, 2 * DMODEL, 1.0, 1, st));
            MDGAT_CUDA_OK(linear(w.Hd, LDHID, 2 * DMODEL, nullptr, 0, 0, Wt + lo.w2, 2 * DMODEL, Wt + lo.b2, w.X, LDX, w.X, LDX, R, DMODEL, 1.0, 0, st));
        }
    }
    prof_mark(ST_GEMM, st);
    // final_proj (mdgat.py:397) and scores = mdesc0^T mdesc1 / sqrt(128) (:430-431) into the couplings
    MDGAT_CUDA_OK(linear(w.X, LDX, DMODEL, nullptr, 0, 0, Wt + lay.wf, DMODEL, Wt + lay.bf, nullptr, 0, w.MD, LDX, R, DMODEL, 1.0, 0, st));
    MDGAT_CUDA_OK(gemm_nt(w.MD, LDX, (long long)N * LDX, w.MD + (size_t)R0 * LDX, LDX, (long long)M * LDX,
                          w.C, M + 1, (long long)(N + 1) * (M + 1), N, M, DMODEL, B, 1.0 / sqrt((double)DMODEL), st));
    prof_mark(ST_SINKHORN, st);
    MDGAT_CUDA_OK(launch_fill_dustbin(w.C, Wt + lay.bin, B, N, M, st));
    MDGAT_CUDA_OK(launch_sinkhorn_fused(w.C, w.u, w.v, w.skscratch, B, N, M, cfg->sinkhorn_iters, st, cfg->sinkhorn_k32 != 0));
    prof_mark(ST_MATCH, st);

    MatchParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.C = w.C; mp.u = w.u; mp.v = w.v; mp.B = B; mp.N = N; mp.M = M;
    mp.match_mode = cfg->match_mode; mp.mutual_check = cfg->mutual_check; mp.match_threshold = cfg->match_threshold;
    mp.loss_mode = cfg->loss_mode; mp.gamma = cfg->triplet_gamma; mp.gt0 = in->d_gt0; mp.gt1 = in->d_gt1;
    mp.matches0 = out->d_matches0; mp.matches1 = out->d_matches1; mp.ms0 = out->d_mscores0; mp.ms1 = out->d_mscores1;
    mp.loss = out->d_loss; mp.nvalid0 = out->d_nvalid0; mp.Z = cfg->write_Z ? out->d_Z : nullptr;
    mp.scratch = w.mscratch; mp.bad = w.bad;
    MDGAT_CUDA_OK(launch_match_extract(mp, st));
    prof_mark(ST_COUNT, st);
    return MDGAT_OK;
}

int mdgat_linear_f64(const double* d_X0, int ldx0, int K0, const double* d_X1, int ldx1, int K1,
                     const double* d_W, int ldw, const double* d_bias, const double* d_Res, int ldres,
                     double* d_Y, int ldy, int R, int Nout, double scale, int relu, void* stream) {
    MDGAT_REQUIRE(K0 > 0 && (ldx0 % 2) == 0 && (ldw % 2) == 0 && (K0 % 2) == 0 && (K1 % 2) == 0,
                  "mdgat_linear_f64: K0, K1, ldx, ldw must be even (16-byte staging)");
    MDGAT_REQUIRE(d_X1 == nullptr || ((K0 % 32) == 0 && (ldx1 % 2) == 0), "mdgat_linear_f64: with a second input K0 must be a multiple of 32");
    MDGAT_CUDA_OK(linear(d_X0, ldx0, K0, d_X1, ldx1, d_X1 ? K1 : 0, d_W, ldw, d_bias, d_Res, ldres, d_Y, ldy, R, Nout,
                         scale, relu, reinterpret_cast<cudaStream_t>(stream), 0));
    return MDGAT_OK;
}

size_t mdgat_linear_i8_scratch_bytes(int R, int K, int slices) { return ozaki_slices_bytes(R, K, slices) + (size_t)((R + 127) / 128 * 128) * (K / 128) * sizeof(double); }

int mdgat_linear_i8(const double* d_X0, int ldx0, int K0, const double* d_X1, int ldx1, int K1,
                    const void* d_Wslices, const double* d_colscale, const double* d_bias, const double* d_Res, int ldres,
                    double* d_Y, int ldy, int R, int Nout, int relu, int slices, void* d_scratch, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int K = K0 + (d_X1 ? K1 : 0);
    MDGAT_REQUIRE(slices >= 4 && slices <= 7, "mdgat_linear_i8: 4..7 slices");
    MDGAT_REQUIRE((K0 % 128) == 0 && (K == 128 || K == 256) && (Nout % 32) == 0 && Nout <= 384, "mdgat_linear_i8: K0 multiple of 128, K in {128,256}, Nout multiple of 32 and <= 384");
    MDGAT_REQUIRE((ldx0 % 2) == 0 && (ldy % 2) == 0, "mdgat_linear_i8: even leading dimensions");
    MDGAT_REQUIRE(!(relu && d_Res == d_Y && d_Res && K > 128), "mdgat_linear_i8: in-place residual with ReLU needs K == 128");
    double* rs = reinterpret_cast<double*>(d_scratch);
    int8_t* xs = reinterpret_cast<int8_t*>(rs + (size_t)((R + 127) / 128 * 128) * (K / 128));
    MDGAT_CUDA_OK(launch_slice_rows(d_X0, ldx0, K0, d_X1, ldx1, d_X1 ? K1 : 0, R, slices, xs, rs, st));
    OzGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.Xs[0] = xs; a.rowscale[0] = rs; a.Xs[1] = xs + ozaki_slices_bytes(R, 128, slices); a.rowscale[1] = rs + (size_t)((R + 127) / 128 * 128);
    a.Ws = reinterpret_cast<const int8_t*>(d_Wslices); a.colscale = d_colscale; a.bias = d_bias;
    a.Res = d_Res; a.ldres = ldres; a.Y = d_Y; a.ldy = ldy; a.R = R; a.Nout = Nout; a.K = K; a.relu = relu; a.epi = EPI_PLAIN;
    MDGAT_CUDA_OK(launch_ozaki_gemm(a, slices, st));
    return MDGAT_OK;
}

int mdgat_gemm_nt_f64(const double* d_X, int ldx, long long sX, const double* d_W, int ldw, long long sW,
                      double* d_Y, int ldy, long long sY, int R, int Nout, int K, int batch, double scale, void* stream) {
    MDGAT_REQUIRE((ldx % 2) == 0 && (ldw % 2) == 0 && (K % 2) == 0 && (sX % 2) == 0 && (sW % 2) == 0,
                  "mdgat_gemm_nt_f64: K, ld and batch strides must be even (16-byte staging)");
    MDGAT_CUDA_OK(gemm_nt(d_X, ldx, sX, d_W, ldw, sW, d_Y, ldy, sY, R, Nout, K, batch, scale, reinterpret_cast<cudaStream_t>(stream)));
    return MDGAT_OK;
}

size_t mdgat_encode_scratch_doubles(int R) { return (size_t)R * (4 + 36 + LDX + LDX + LDX); }

int mdgat_encode(const mdgat_forward_in* in, int B, int N, int M, int in_dtype, int score_dtype,
                 const double* d_weights, double* d_X, double* d_tmp, void* stream) {
    const size_t R = (size_t)B * N + (size_t)B * M;
    double* Xk = d_tmp; double* Xd = Xk + R * 4; double* T0 = Xd + R * 36; double* T1 = T0 + R * LDX; double* T2 = T1 + R * LDX;
    const BlobLayout lay(1);      // encoder offsets do not depend on L
    MDGAT_CUDA_OK(encode(in, B, N, M, in_dtype, score_dtype, d_weights, lay, d_X, Xk, Xd, T0, T1, T2, LDX,
                         reinterpret_cast<cudaStream_t>(stream)));
    return MDGAT_OK;
}

size_t mdgat_attention_f64_scratch_doubles(int B, int N, int M) {
    if (B < 1 || N < 1 || M < 1) return 0;
    AttnSides ps;
    memset(&ps, 0, sizeof(ps));
    ps.N[0] = N; ps.M[0] = M;
    const size_t dense = (size_t)B * HEADS * N * M, ring = topk_fused_ring_need(ps, B, 1);
    return dense > ring ? dense : ring;
}

int mdgat_attention_f64(const double* d_Q, const double* d_K, const double* d_V, double* d_Out, int ldo,
                        int B, int N, int M, int topk, double* d_logits, void* stream) {
    MDGAT_REQUIRE(topk <= M, "selected index k out of range (k=%d, M=%d)", topk, M);
    MDGAT_REQUIRE(topk <= 0 || (d_logits != nullptr && M <= 2048), "mdgat_attention_f64: top-k needs a logits scratch and M <= 2048");
    MDGAT_REQUIRE((ldo % 2) == 0, "mdgat_attention_f64: ldo must be even");
    AttnSides ps;
    memset(&ps, 0, sizeof(ps));
    ps.Q[0] = d_Q; ps.K[0] = d_K; ps.V[0] = d_V; ps.Out[0] = d_Out; ps.N[0] = N; ps.M[0] = M;
    MDGAT_CUDA_OK(attention_layer(ps, B, 1, ldo, topk, d_logits, reinterpret_cast<cudaStream_t>(stream), nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                                  mdgat_attention_f64_scratch_doubles(B, N, M)));
    return MDGAT_OK;
}

size_t mdgat_attention_i8_scratch_bytes(int B, int N, int M) {
    // digit planes of both sets (sized for 7 planes) + the top-k scratch: threshold, maximum (doubles), last tied column (int) per (b, h, row)
    return attn_i8_side_bytes(B, N, 7) + attn_i8_side_bytes(B, M, 7) + (size_t)B * HEADS * N * 20 + 256;
}

int mdgat_attention_i8(const double* d_Q, const double* d_K, const double* d_V, double* d_Out, int ldo,
                       int B, int N, int M, int topk, double* d_logits, void* d_scratch, int slices, int p_slices,
                       void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int AS = slices > 0 ? slices : 7, ASP = p_slices > 0 ? p_slices : AS - 1;
    MDGAT_REQUIRE(attn_planes_ok(AS, ASP), "mdgat_attention_i8: unsupported digit planes (slices %d, p_slices %d)", AS, ASP);
    MDGAT_REQUIRE(topk <= M, "selected index k out of range (k=%d, M=%d)", topk, M);
    MDGAT_REQUIRE(topk <= 0 || (d_logits != nullptr && M <= 2048), "mdgat_attention_i8: top-k needs a logits scratch and M <= 2048");
    MDGAT_REQUIRE((ldo % 2) == 0 && d_scratch != nullptr, "mdgat_attention_i8: ldo must be even, scratch required");
    MDGAT_REQUIRE(attn_i8_supported(N, M), "mdgat_attention_i8: M = %d source keypoints exceed the shared-memory budget", M);
    // query planes from Q of the N-point set, source planes from K and V of the M-point set
    AttnI8Side qs = attn_i8_carve(d_scratch, B, N, AS);
    AttnI8Side ks = attn_i8_carve(reinterpret_cast<char*>(d_scratch) + attn_i8_side_bytes(B, N, AS), B, M, AS);
    MDGAT_CUDA_OK(launch_attn_i8_slice(d_Q, nullptr, nullptr, qs, B, st));
    MDGAT_CUDA_OK(launch_attn_i8_slice(nullptr, d_K, d_V, ks, B, st));
    AttnSides ps;
    memset(&ps, 0, sizeof(ps));
    ps.Q[0] = d_Q; ps.K[0] = d_K; ps.V[0] = d_V; ps.Out[0] = d_Out; ps.N[0] = N; ps.M[0] = M;
    char* tkb = reinterpret_cast<char*>(d_scratch) + attn_i8_side_bytes(B, N, 7) + attn_i8_side_bytes(B, M, 7);
    const size_t rows = (size_t)B * HEADS * N;
    const TopKScratch tks = {reinterpret_cast<double*>(tkb), reinterpret_cast<double*>(tkb) + rows, reinterpret_cast<int*>(tkb + rows * 16)};
    MDGAT_CUDA_OK(attention_layer(ps, B, 1, ldo, topk, d_logits, st, &qs, &ks, ASP, &tks));
    return MDGAT_OK;
}

size_t mdgat_sinkhorn_scratch_doubles(int B, int N, int M) { return sinkhorn_scratch_doubles(B, N, M); }

int mdgat_sinkhorn_read_status(const double* d_scratch, int B, int N, int M, int* h_flags, int* h_iters) {
    const size_t ldk = (size_t)((M + 1) & ~1);
    const int* d = reinterpret_cast<const int*>(d_scratch + (size_t)B * (N + 1) * ldk);
    MDGAT_CUDA_OK(cudaMemcpy(h_flags, d, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost));
    MDGAT_CUDA_OK(cudaMemcpy(h_iters, d + B, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost));
    return MDGAT_OK;
}

int mdgat_forward_sinkhorn_status(const mdgat_forward_cfg* cfg, const void* d_workspace, int* h_flags, int* h_iters) {
    MDGAT_REQUIRE(cfg && d_workspace && h_flags && h_iters, "mdgat_forward_sinkhorn_status: bad arguments");
    bool need = false;
    for (int i = 0; i < 2 * cfg->L; ++i) need = need || (cfg->layer_k && cfg->layer_k[i] > 0);
    // the Sinkhorn scratch lies before every engine-dependent part of the workspace
    const Workspace w = carve(reinterpret_cast<char*>(const_cast<void*>(d_workspace)), cfg->B, cfg->N, cfg->M, need);
    return mdgat_sinkhorn_read_status(w.skscratch, cfg->B, cfg->N, cfg->M, h_flags, h_iters);
}

size_t mdgat_attention_backward_scratch_doubles(int B, int N, int M, int topk) { return attention_bwd_scratch_doubles(B, N, M, topk); }

int mdgat_attention_backward_f64(const double* d_Q, const double* d_K, const double* d_V, const double* d_O, const double* d_dO,
                                 double* d_dQ, double* d_dK, double* d_dV, int B, int N, int M, int topk, double* d_scratch,
                                 void* stream) {
    MDGAT_REQUIRE(B > 0 && N > 0 && M > 0 && d_Q && d_K && d_V && d_O && d_dO && d_dQ && d_dK && d_dV && d_scratch, "mdgat_attention_backward_f64: bad arguments");
    MDGAT_REQUIRE(topk <= M, "selected index k out of range (k=%d, M=%d)", topk, M);
    MDGAT_REQUIRE(topk <= 0 || M <= 2048, "mdgat_attention_backward_f64: top-k supports at most 2048 sources (M=%d)", M);
    MDGAT_CUDA_OK(launch_attention_backward(d_Q, d_K, d_V, d_O, d_dO, d_dQ, d_dK, d_dV, B, N, M, topk > 0 ? topk : 0, d_scratch,
                                            reinterpret_cast<cudaStream_t>(stream)));
    return MDGAT_OK;
}

size_t mdgat_sinkhorn_backward_scratch_doubles(int B, int N, int M, int iters) { return sinkhorn_bwd_scratch_doubles(B, N, M, iters); }

int mdgat_sinkhorn_backward_f64(const double* d_couplings, const double* d_gZ, double* d_gcouplings, int B, int N, int M,
                                int iters, double* d_scratch, int* h_ill_conditioned, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MDGAT_REQUIRE(B > 0 && N > 0 && M > 0 && iters >= 0 && d_couplings && d_gZ && d_gcouplings && d_scratch, "mdgat_sinkhorn_backward_f64: bad arguments");
    MDGAT_CUDA_OK(launch_sinkhorn_backward(d_couplings, d_gZ, d_gcouplings, d_scratch, B, N, M, iters, st));
    if (h_ill_conditioned) {
        *h_ill_conditioned = 0;
        if (iters > 0) {
            const int* flag = reinterpret_cast<const int*>(d_scratch + (mdgat_sinkhorn_backward_scratch_doubles(B, N, M, iters) - 2));
            MDGAT_CUDA_OK(cudaMemcpyAsync(h_ill_conditioned, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
            MDGAT_CUDA_OK(cudaStreamSynchronize(st));
        }
    }
    return MDGAT_OK;
}

int mdgat_sinkhorn_f64(double* d_couplings, const double* d_bin_score, double* d_u, double* d_v,
                       int B, int N, int M, int iters, double* d_scratch, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MDGAT_REQUIRE(B > 0 && N > 0 && M > 0 && iters >= 0, "mdgat_sinkhorn_f64: bad shape");
    MDGAT_CUDA_OK(launch_fill_dustbin(d_couplings, d_bin_score, B, N, M, st));
    if (d_scratch) MDGAT_CUDA_OK(launch_sinkhorn_fused(d_couplings, d_u, d_v, d_scratch, B, N, M, iters, st));
    else MDGAT_CUDA_OK(launch_sinkhorn(d_couplings, d_u, d_v, B, N, M, iters, st));
    return MDGAT_OK;
}

int mdgat_sinkhorn_f64_k32(double* d_couplings, const double* d_bin_score, double* d_u, double* d_v,
                       int B, int N, int M, int iters, double* d_scratch, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MDGAT_REQUIRE(B > 0 && N > 0 && M > 0 && iters >= 0, "mdgat_sinkhorn_f64_k32: bad shape");
    MDGAT_CUDA_OK(launch_fill_dustbin(d_couplings, d_bin_score, B, N, M, st));
    MDGAT_REQUIRE(d_scratch != nullptr, "mdgat_sinkhorn_f64_k32: the fused kernel needs its scratch (mdgat_sinkhorn_scratch_doubles)");
    MDGAT_CUDA_OK(launch_sinkhorn_fused(d_couplings, d_u, d_v, d_scratch, B, N, M, iters, st, true));
    return MDGAT_OK;
}

size_t mdgat_match_scratch_doubles(int B, int N, int M) { return 4 * ((size_t)B * N + (size_t)B * M) + 8; }

int mdgat_match_extract(const double* d_couplings, const double* d_u, const double* d_v, int B, int N, int M,
                        int match_mode, int mutual_check, double match_threshold, int loss_mode, double gamma,
                        const int16_t* d_gt0, const int16_t* d_gt1, const mdgat_forward_out* out,
                        double* d_scratch, void* stream) {
    MDGAT_REQUIRE(loss_mode >= MDGAT_LOSS_NONE && loss_mode <= MDGAT_LOSS_SUPERGLUE, "mdgat_match_extract: unknown loss_mode %d", loss_mode);
    MDGAT_REQUIRE(loss_mode != MDGAT_LOSS_TRIPLET || (d_gt0 && d_gt1 && N == M && match_mode == MDGAT_MATCH_DUSTBIN),
                  "mdgat_match_extract: triplet loss needs gt, N == M and the dustbin variant");
    MDGAT_REQUIRE(loss_mode != MDGAT_LOSS_GAP || (d_gt0 && d_gt1 && match_mode == MDGAT_MATCH_DUSTBIN), "mdgat_match_extract: gap loss needs gt and the dustbin variant");
    MDGAT_REQUIRE(loss_mode != MDGAT_LOSS_SUPERGLUE || (d_gt0 && d_gt1 && N == M), "mdgat_match_extract: superglue loss needs gt and N == M");
    MatchParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.C = d_couplings; mp.u = d_u; mp.v = d_v; mp.B = B; mp.N = N; mp.M = M;
    mp.match_mode = match_mode; mp.mutual_check = mutual_check; mp.match_threshold = match_threshold;
    mp.loss_mode = loss_mode; mp.gamma = gamma; mp.gt0 = d_gt0; mp.gt1 = d_gt1;
    mp.matches0 = out->d_matches0; mp.matches1 = out->d_matches1; mp.ms0 = out->d_mscores0; mp.ms1 = out->d_mscores1;
    mp.loss = out->d_loss; mp.nvalid0 = out->d_nvalid0; mp.Z = out->d_Z; mp.scratch = d_scratch;
    MDGAT_CUDA_OK(launch_match_extract(mp, reinterpret_cast<cudaStream_t>(stream)));
    return MDGAT_OK;
}

int mdgat_knn(const double* d_x, const double* d_src, int64_t* d_idx, int B, int n, int m, int k, void* stream) {
    MDGAT_REQUIRE(k > 0 && k <= m, "selected index k out of range (k=%d, m=%d)", k, m);
    MDGAT_REQUIRE(m <= 2048, "mdgat_knn supports at most 2048 source points (m=%d)", m);
    MDGAT_CUDA_OK(launch_knn(d_x, d_src, d_idx, B, n, m, k, reinterpret_cast<cudaStream_t>(stream)));
    return MDGAT_OK;
}

long long mdgat_launch_count(void) { return g_launches.load(); }
void mdgat_launch_count_add(long long n) { g_launches.fetch_add(n); }

int mdgat_debug_trace(void* d_buf) { mdgat::g_trace_dev = reinterpret_cast<long long*>(d_buf); return MDGAT_OK; }

int mdgat_debug_flags(int flags) { mdgat::g_debug_flags = flags; return MDGAT_OK; }

int mdgat_profile_enable(int on) {
    g_prof.on = on != 0;
    g_prof.used = 0; g_prof.stage.clear(); g_prof.launches_at.clear();
    for (int i = 0; i < ST_COUNT; ++i) { g_prof.ms[i] = 0; g_prof.launches[i] = 0; g_prof.segments[i] = 0; }
    return MDGAT_OK;
}

int mdgat_profile_collect(double* ms, long long* launches, long long* segments, int n) {
    MDGAT_REQUIRE(n >= ST_COUNT, "mdgat_profile_collect: need room for %d stages", (int)ST_COUNT);
    if (g_prof.used > 0) MDGAT_CUDA_OK(cudaEventSynchronize(g_prof.ev[g_prof.used - 1]));
    for (size_t i = 0; i + 1 < g_prof.used; ++i) {
        const int s = g_prof.stage[i];
        if (s >= ST_COUNT) continue;                 // gap between two forwards
        float t = 0.f;
        MDGAT_CUDA_OK(cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]));
        g_prof.ms[s] += t;
        g_prof.launches[s] += g_prof.launches_at[i + 1] - g_prof.launches_at[i];
        g_prof.segments[s] += 1;
    }
    g_prof.used = 0; g_prof.stage.clear(); g_prof.launches_at.clear();
    for (int i = 0; i < ST_COUNT; ++i) { ms[i] = g_prof.ms[i]; launches[i] = g_prof.launches[i]; segments[i] = g_prof.segments[i]; }
    return MDGAT_OK;
}

int mdgat_prepare_pairs(const double* d_kp1, const double* d_kp2, const double* d_pose1, const double* d_pose2,
                        const double* d_T_cam0_velo, int calib_per_pair, int B, int N, int M, double threshold,
                        int mutual_check, int16_t* d_match1, int16_t* d_match2, double* d_T_gt, int* d_rep, void* stream) {
    MDGAT_REQUIRE(B >= 0 && N > 0 && M > 0 && N < 32768 && M < 32768, "mdgat_prepare_pairs: bad shape (int16 match indices)");
    MDGAT_REQUIRE(prepare_pairs_smem(N, M) <= 227 * 1024, "mdgat_prepare_pairs: N + M = %d keypoints do not fit in shared memory", N + M);
    MDGAT_CUDA_OK(launch_prepare_pairs(d_kp1, d_kp2, d_pose1, d_pose2, d_T_cam0_velo, calib_per_pair, B, N, M, threshold,
                                       mutual_check, d_match1, d_match2, d_T_gt, d_rep, reinterpret_cast<cudaStream_t>(stream)));
    return MDGAT_OK;
}

int mdgat_register_pairs(const void* d_kpts0, const void* d_kpts1, int kp_dtype, const int64_t* d_matches0,
                         const int16_t* d_gt0, const double* d_T_gt, int B, int N, int M,
                         double* d_T, double* d_stats, void* stream) {
    MDGAT_REQUIRE(B >= 0 && N > 0 && M > 0 && d_T && d_stats, "mdgat_register_pairs: bad arguments");
    MDGAT_CUDA_OK(launch_register_pairs(d_kpts0, d_kpts1, kp_dtype, d_matches0, d_gt0, d_T_gt, B, N, M, d_T, d_stats,
                                        reinterpret_cast<cudaStream_t>(stream)));
    return MDGAT_OK;
}

int mdgat_measure_fp64_mixed(double* tflops_dmma, double* tflops_dfma) {
    MDGAT_CUDA_OK(measure_fp64_mixed(tflops_dmma, tflops_dfma));
    MDGAT_CUDA_OK(measure_dmma_tiled(tflops_dfma + 1));
    return MDGAT_OK;
}

int mdgat_measure_i8_peak(double* tops) {
    MDGAT_CUDA_OK(measure_i8_peak(tops));
    return MDGAT_OK;
}

int mdgat_measure_fp64_peak(double* tflops_dmma, double* tflops_dfma) {
    MDGAT_CUDA_OK(measure_fp64_peak(tflops_dmma, tflops_dfma));
    return MDGAT_OK;
}

}  // extern "C"
