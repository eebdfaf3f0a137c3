// Shared device helpers for the sm_100a kernels of mdgat-matcher_b200.
//
// Arithmetic policy (DESIGN.md "Precision"): the reference runs in float64 end to end
// (/root/reference/test.py:193) and its top-k layers make the result discontinuous in the
// logits, so everything that feeds a dynamic layer is computed in float64. tcgen05.mma has
// no f64 kind; the f64 tensor path on sm_100a is mma.sync.m8n8k4 (SASS: DMMA.8x8x4).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "exp_table.h"

#define DEVINL __device__ __forceinline__

namespace mdgat {

constexpr int HEADS = 4;        // AttentionalPropagation(feature_dim, 4), mdgat.py:255
constexpr int HDIM = 32;        // 128 / 4
constexpr int DMODEL = 128;
constexpr int LDH_QK = 36;      // row stride (doubles) of head-major Q and K buffers
constexpr int LDH_V = 34;       // row stride (doubles) of the head-major V buffer
constexpr int LDX = 132;        // row stride (doubles) of 128-wide activation buffers

// D(8x8) += A(8x4, row) * B(4x8, col). Lane t holds A[t/4][t%4], B[t%4][t/4],
// C[t/4][2*(t%4) + {0,1}].
DEVINL void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 16-byte async copy global -> shared (LDGSTS); bytes beyond src_bytes are zero-filled.
DEVINL void cp_async16(void* smem, const void* gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem), "r"(sz));
}
DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

// ---- mbarrier + bulk (TMA engine) copies: 1-D cp.async.bulk global -> shared::cta ----
DEVINL void mbar_init(uint64_t* bar, unsigned count) {
    unsigned s = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(s), "r"(count));
}
DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::); }
DEVINL void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(s), "r"(bytes) : "memory");
}
DEVINL void mbar_arrive(uint64_t* bar) {
    unsigned s = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(s) : "memory");
}
// suspend-time hint of try_wait: the waiting warp sleeps in hardware until the phase completes (wake-up ~60 cycles after the
// arrive) or the hint expires; without a hint the default time-out is short and a warp waiting for the tensor core re-issues
// TRYWAIT + BRA half a dozen times per wait, taking issue slots from the three other warps of its scheduler
#ifndef MBAR_SUSPEND_HINT
#define MBAR_SUSPEND_HINT 0x989680u
#endif
DEVINL bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned s = (unsigned)__cvta_generic_to_shared(bar), ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(s), "r"(parity), "r"(MBAR_SUSPEND_HINT) : "memory");
    return ok != 0;
}
DEVINL void mbar_wait(uint64_t* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// bytes must be a multiple of 16; both addresses 16-byte aligned.
DEVINL void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(d), "l"(gmem), "r"(bytes), "r"(b) : "memory");
}

// One lane of a converged warp. Single-thread roles (TMA producer, tcgen05.mma issuer) must be entered through this and not
// through `lane == 0`: with a lane test the compiler wraps every uniform-datapath instruction (UTCIMMA, UBLKCP, ..) in
// its own ELECT / BRA.U.ANY retry loop (11 instructions between two MMAs instead of 2-4).
DEVINL bool elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

// ---- timeline probe (debug): one CTA records (tag, clock64) pairs per role into a global buffer set with
// mdgat_debug_trace() and handed to the kernels as a parameter; a null buffer (the default) costs one predicate per probe.
constexpr int TRACE_CAP = 1024;                 // records per role
struct Tracer {
    long long* base; int n;
    DEVINL void init(long long* b, int role, bool cta_selected) {
        base = (b != nullptr && cta_selected) ? b + (size_t)role * (2 * TRACE_CAP + 2) : nullptr; n = 0;
    }
    DEVINL void mark(int tag) {
        if (base != nullptr && n < TRACE_CAP) { base[2 + 2 * n] = tag; base[3 + 2 * n] = clock64(); ++n; base[0] = n; }
    }
};

// ---- programmatic dependent launch (PDL). A kernel launched through launch_pdl() may become resident while the kernel
// before it in the stream is still draining: everything up to pdl_wait() (barrier / TMEM set-up, tables from constant
// data) overlaps that tail. pdl_wait() returns once the preceding kernel has completed and its writes are visible; no
// global memory the forward pass produces may be read or written before it. pdl_trigger() lets the NEXT kernel in the
// stream start its own prologue; it is issued right after the wait, so at most one grid is parked at a time.
DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

DEVINL double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
DEVINL double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

DEVINL double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, shfl_xor_d(v, o));
    return v;
}
DEVINL double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

// exp(x) for x <= 0 (softmax arguments after max subtraction), about 1-2 ulp.
// x = (64 m + j) ln2/64 + r, |r| <= ln2/128: exp(x) = 2^m * 2^(j/64) * e^r with a 64-entry
// table (in shared memory: lanes index it at random) and a degree-5 polynomial evaluated in
// Estrin form. 11 FP64-pipe instructions with a 9-deep dependency chain, against 17 / 17 for
// libm exp(): the softmax exponentials share the FP64 pipe with the DMMAs around them.
// Results below 2^-1022 are flushed to zero.
__device__ const double g_exp2_table[64] = MDGAT_EXP2_TABLE;

DEVINL void exp_table_to_shared(double* tbl) {
    if (threadIdx.x < 64) tbl[threadIdx.x] = g_exp2_table[threadIdx.x];
}

DEVINL double exp_fast_neg(double x, const double* tbl) {
    const double xc = fmax(x, -745.0);
    const double MAGIC = 6755399441055744.0;                // 1.5 * 2^52: rint() in the low mantissa bits
    const double tn = fma(xc, EXP_INV_LN2_64, MAGIC);
    const int n = __double2loint(tn);
    const double nd = tn - MAGIC;
    double r = fma(nd, -EXP_LN2_64_HI, xc);
    r = fma(nd, -EXP_LN2_64_LO, r);
    const double r2 = r * r;
    const double a = fma(r, 1.0 / 6.0, 0.5);
    const double b = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double c = fma(r2, b, a);
    const double p = fma(r2, c, r);                         // e^r - 1
    const double t = tbl[n & 63];
    const double y = fma(t, p, t);                          // in [1, 2)
    const int m = n >> 6;
    const double ys = __hiloint2double(__double2hiint(y) + (m << 20), __double2loint(y));
    return xc < -708.0 ? 0.0 : ys;
}

// exp(x) for x <= 0 with an ABSOLUTE error far below 2^-47, for probabilities that are cut into 47-bit fixed-point digits
// (attention_i8.cu): 256-entry table, |r| <= ln2/512 so a degree-4 polynomial reaches 4e-17, and one correctly rounded
// ln2/256 is enough under the FMA (its error scales with |x| e^x <= 1/e: 3e-17 absolute). 9 FP64 instructions against
// 11 of exp_fast_neg.
__device__ const double g_exp2_table256[256] = MDGAT_EXP2_TABLE256;
DEVINL void exp_table256_to_shared(double* tbl) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tbl[i] = g_exp2_table256[i];
}
DEVINL double exp_neg_abs47(double x, const double* tbl) {
    const double xc = fmax(x, -745.0);
    const double MAGIC = 6755399441055744.0;                // 1.5 * 2^52: rint() in the low mantissa bits
    const double tn = fma(xc, EXP_INV_LN2_256, MAGIC);
    const int n = __double2loint(tn);
    const double nd = tn - MAGIC;
    const double r = fma(nd, -EXP_LN2_256, xc);
    const double r2 = r * r;
    double q = fma(r, 1.0 / 24.0, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    const double p = fma(r2, q, r);                         // e^r - 1
    const double t = tbl[n & 255];
    const double y = fma(t, p, t);                          // in [1, 2)
    const int m = n >> 8;
    const double ys = __hiloint2double(__double2hiint(y) + (m << 20), __double2loint(y));
    return xc < -708.0 ? 0.0 : ys;
}

}  // namespace mdgat

// ---- host-side error plumbing shared by the C ABI translation units ----
namespace mdgat_host {
void set_error(const char* fmt, ...);
}

// ---- host side of PDL: launch with the programmatic-stream-serialization attribute (MDGAT_PDL=0 switches it off; the
// kernels' griddepcontrol instructions are no-ops then). Only kernels that call pdl_wait() may be launched this way.
namespace mdgat {
bool pdl_enabled();                                        // capi.cu
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
}  // namespace mdgat
#define MDGAT_CUDA_OK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            mdgat_host::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                  __FILE__, __LINE__);                                   \
            return MDGAT_ERR_CUDA;                                                       \
        }                                                                                \
    } while (0)
#define MDGAT_REQUIRE(cond, ...)                     \
    do {                                             \
        if (!(cond)) {                               \
            mdgat_host::set_error(__VA_ARGS__);      \
            return MDGAT_ERR_INVALID;                \
        }                                            \
    } while (0)
