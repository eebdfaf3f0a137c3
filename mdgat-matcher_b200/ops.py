"""Torch-tensor front ends of the individual C-ABI operators (include/mdgat_b200.h).

These are the same kernels mdgat_forward() launches; they exist so that each block of the
reference can be checked on its own. Inputs must be CUDA tensors.
"""
import ctypes
import math

import torch

from . import _capi
from ._capi import LDX, LDH_QK, LDH_V


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _need_cuda(t):
    if not t.is_cuda:
        raise RuntimeError('mdgat-matcher_b200 operators need CUDA tensors (no CPU fallback)')


def linear(x, w, bias=None, relu=False, residual=None, x2=None, scale=1.0):
    """y = act(scale * [x | x2] w^T + bias) + residual; x (R,K0), w (Nout,K) float64."""
    _need_cuda(x)
    x = x.double().contiguous()
    w = w.double().contiguous()
    R, K0 = x.shape
    K1 = 0
    if x2 is not None:
        x2 = x2.double().contiguous()
        K1 = x2.shape[1]
    nout = w.shape[0]
    y = torch.empty((R, nout), dtype=torch.float64, device=x.device)
    b = bias.double().contiguous() if bias is not None else None
    r = residual.double().contiguous() if residual is not None else None
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib.mdgat_linear_f64(
            x.data_ptr(), x.stride(0), K0, x2.data_ptr() if x2 is not None else None,
            x2.stride(0) if x2 is not None else 0, K1, w.data_ptr(), w.stride(0),
            b.data_ptr() if b is not None else None, r.data_ptr() if r is not None else None,
            r.stride(0) if r is not None else 0, y.data_ptr(), y.stride(0), R, nout, float(scale), int(relu),
            _stream(x.device)))
    return y


def gemm_nt(x, w, scale=1.0):
    """Batched y[z] = scale * x[z] w[z]^T; x (Z,R,K), w (Z,Nout,K) float64."""
    _need_cuda(x)
    x = x.double().contiguous()
    w = w.double().contiguous()
    Z, R, K = x.shape
    nout = w.shape[1]
    y = torch.empty((Z, R, nout), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib.mdgat_gemm_nt_f64(x.data_ptr(), K, R * K, w.data_ptr(), K, nout * K,
                                                y.data_ptr(), nout, R * nout, R, nout, K, Z, float(scale),
                                                _stream(x.device)))
    return y


def to_head_major(t, ld):
    """(B, 128, N) reference-layout q/k/v (channel c = d*4 + h, mdgat.py:227) -> (B,4,N,ld) padded."""
    b, c, n = t.shape
    x = t.double().reshape(b, 32, 4, n).permute(0, 2, 3, 1)          # (B, h, N, d)
    out = torch.zeros((b, 4, n, ld), dtype=torch.float64, device=t.device)
    out[..., :32] = x
    return out.contiguous()


def attention(q, k, v, topk=None, engine='dmma', slices=7, p_slices=0):
    """q (B,128,N), k/v (B,128,M) in the reference's channel layout. Returns the message
    (B,128,N) in the reference layout (what attention()/dynamic_attention() return after
    .view(B, 128, N), mdgat.py:229-237 before the merge conv)."""
    _need_cuda(q)
    B, _, N = q.shape
    M = k.shape[2]
    qh, kh, vh = to_head_major(q, LDH_QK), to_head_major(k, LDH_QK), to_head_major(v, LDH_V)
    out = torch.empty((B * N, LDX), dtype=torch.float64, device=q.device)
    kk = 0 if topk is None else int(topk)
    logits = torch.empty(max(B * 4 * N * M, _capi.lib.mdgat_attention_f64_scratch_doubles(B, N, M)), dtype=torch.float64, device=q.device) if kk > 0 else None
    with torch.cuda.device(q.device):
        if engine == 'tcgen05_i8':
            scratch = torch.empty(_capi.lib.mdgat_attention_i8_scratch_bytes(B, N, M), dtype=torch.uint8, device=q.device)
            _capi.check(_capi.lib.mdgat_attention_i8(qh.data_ptr(), kh.data_ptr(), vh.data_ptr(), out.data_ptr(), LDX,
                                                     B, N, M, kk, logits.data_ptr() if logits is not None else None,
                                                     scratch.data_ptr(), int(slices), int(p_slices), _stream(q.device)))
        elif engine == 'dmma':
            _capi.check(_capi.lib.mdgat_attention_f64(qh.data_ptr(), kh.data_ptr(), vh.data_ptr(), out.data_ptr(), LDX,
                                                      B, N, M, kk, logits.data_ptr() if logits is not None else None,
                                                      _stream(q.device)))
        else:
            raise ValueError("engine must be 'dmma' or 'tcgen05_i8'")
    msg = out[:, :128].reshape(B, N, 4, 32)                        # (B, N, h, d)
    return msg.permute(0, 3, 2, 1).reshape(B, 128, N).contiguous()  # channel c = d*4 + h


class AttentionFn(torch.autograd.Function):
    """attention() / dynamic_attention() of mdgat.py:190-210 for the training path: forward = the CUDA kernels (float64 DMMA
    flash attention, exact top-k), backward = the hand-written tile-recompute kernels of csrc/attention_bwd.cu. Only q, k, v
    and the message are kept for backward; autograd through the reference formulation keeps the (B,4,N,M) probabilities."""

    @staticmethod
    def forward(ctx, q, k, v, topk):
        _need_cuda(q)
        B, _, N = q.shape
        M = k.shape[2]
        kk = int(topk) if topk else 0
        qh, kh, vh = to_head_major(q.detach(), LDH_QK), to_head_major(k.detach(), LDH_QK), to_head_major(v.detach(), LDH_V)
        out = torch.empty((B * N, LDX), dtype=torch.float64, device=q.device)
        logits = torch.empty(max(B * 4 * N * M, _capi.lib.mdgat_attention_f64_scratch_doubles(B, N, M)), dtype=torch.float64,
                             device=q.device) if kk > 0 else None
        with torch.cuda.device(q.device):
            _capi.check(_capi.lib.mdgat_attention_f64(qh.data_ptr(), kh.data_ptr(), vh.data_ptr(), out.data_ptr(), LDX, B, N, M, kk,
                                                      logits.data_ptr() if logits is not None else None, _stream(q.device)))
        oh = out[:, :128].reshape(B, N, 4, 32).permute(0, 2, 1, 3).contiguous()        # (B, h, N, d)
        ctx.save_for_backward(qh, kh, vh, oh)
        ctx.kk, ctx.dtypes = kk, (q.dtype, k.dtype, v.dtype)
        return oh.permute(0, 3, 1, 2).reshape(B, 128, N)                               # channel c = d*4 + h

    @staticmethod
    def backward(ctx, g):
        qh, kh, vh, oh = ctx.saved_tensors
        B, _, N, _ = qh.shape
        M = kh.shape[2]
        dev = qh.device
        doh = g.double().reshape(B, 32, 4, N).permute(0, 2, 3, 1).contiguous()         # (B, h, N, d)
        dq = torch.empty((B, 4, N, 32), dtype=torch.float64, device=dev)
        dk = torch.empty((B, 4, M, 32), dtype=torch.float64, device=dev)
        dv = torch.empty((B, 4, M, 32), dtype=torch.float64, device=dev)
        scratch = torch.empty(_capi.lib.mdgat_attention_backward_scratch_doubles(B, N, M, ctx.kk), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _capi.check(_capi.lib.mdgat_attention_backward_f64(qh.data_ptr(), kh.data_ptr(), vh.data_ptr(), oh.data_ptr(), doh.data_ptr(),
                                                               dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, N, M, ctx.kk,
                                                               scratch.data_ptr(), _stream(dev)))
        back = lambda t, n, dt: t.permute(0, 3, 1, 2).reshape(B, 128, n).to(dt)
        return back(dq, N, ctx.dtypes[0]), back(dk, M, ctx.dtypes[1]), back(dv, M, ctx.dtypes[2]), None


def attention_autograd(q, k, v, topk=None):
    """Differentiable message (B,128,N) of attention() / dynamic_attention() on the CUDA kernels (forward and backward)."""
    return AttentionFn.apply(q, k, v, topk)


def sinkhorn(scores, bin_score, iters, fused=True, return_status=False, k32=False):
    """scores (B,N,M) -> (couplings, u, v) with Z = couplings + u[:, :, None] + v[:, None, :] - norm
    = log_optimal_transport(scores, bin_score, iters). fused=False uses one launch per half-iteration; k32=True stores
    the kernel matrix exp(C - rowmax) in float32 (float64 arithmetic), the forward's default ('sweep' precision)."""
    _need_cuda(scores)
    B, N, M = scores.shape
    dev = scores.device
    C = torch.empty((B, N + 1, M + 1), dtype=torch.float64, device=dev)
    C[:, :N, :M] = scores.double()
    u = torch.empty((B, N + 1), dtype=torch.float64, device=dev)
    v = torch.empty((B, M + 1), dtype=torch.float64, device=dev)
    alpha = torch.as_tensor(bin_score, dtype=torch.float64, device=dev).reshape(1).contiguous()
    scratch = torch.empty(_capi.lib.mdgat_sinkhorn_scratch_doubles(B, N, M), dtype=torch.float64, device=dev) if fused else None
    with torch.cuda.device(dev):
        fn = _capi.lib.mdgat_sinkhorn_f64_k32 if (k32 and fused) else _capi.lib.mdgat_sinkhorn_f64
        _capi.check(fn(C.data_ptr(), alpha.data_ptr(), u.data_ptr(), v.data_ptr(),
                       B, N, M, int(iters), scratch.data_ptr() if fused else None, _stream(dev)))
        if return_status and fused:
            fl, it = (ctypes.c_int * B)(), (ctypes.c_int * B)()
            _capi.check(_capi.lib.mdgat_sinkhorn_read_status(scratch.data_ptr(), B, N, M, fl, it))
            return C, u, v, {'fallback': list(fl), 'iterations': list(it)}
    return C, u, v


def sinkhorn_backward(couplings, gZ, iters):
    """dL/d(couplings) (B,N+1,M+1) of Z = log_optimal_transport from dL/dZ, by the hand-written reverse sweep
    (csrc/sinkhorn_bwd.cu); `couplings` must carry the dustbin row / column (what sinkhorn() returns as its first result)."""
    _need_cuda(couplings)
    B, N, M = couplings.shape[0], couplings.shape[1] - 1, couplings.shape[2] - 1
    dev = couplings.device
    C = couplings.double().contiguous()
    G = gZ.double().contiguous()
    out = torch.empty_like(C)
    scratch = torch.empty(_capi.lib.mdgat_sinkhorn_backward_scratch_doubles(B, N, M, int(iters)), dtype=torch.float64, device=dev)
    ill = ctypes.c_int(0)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib.mdgat_sinkhorn_backward_f64(C.data_ptr(), G.data_ptr(), out.data_ptr(), B, N, M, int(iters),
                                                          scratch.data_ptr(), ctypes.byref(ill), _stream(dev)))
    if ill.value:
        raise RuntimeError("mdgat-matcher_b200: a row of the Sinkhorn couplings spans more than 600 -- the scaling-form backward "
                           "cannot represent it; set config['cuda_sinkhorn_backward'] = False to back-propagate through the "
                           "unrolled torch iterations instead")
    return out


class LogOptimalTransportFn(torch.autograd.Function):
    """log_optimal_transport (mdgat.py:279-308) for the training path: forward = the fused float64 Sinkhorn kernel, backward =
    the hand-written reverse sweep. Nothing but the couplings is kept between the two (autograd through the reference's
    unrolled loop retains 2 T tensors of shape (B, N+1, M+1))."""

    @staticmethod
    def forward(ctx, scores, alpha, iters):
        B, N, M = scores.shape
        C, u, v = sinkhorn(scores.detach(), alpha.detach(), iters, fused=True, k32=False)
        ctx.save_for_backward(C)
        ctx.iters = int(iters)
        ctx.in_dtype = scores.dtype
        norm = -math.log(N + M)
        return C + u[:, :, None] + v[:, None, :] - norm

    @staticmethod
    def backward(ctx, gZ):
        C, = ctx.saved_tensors
        gC = sinkhorn_backward(C, gZ, ctx.iters)
        N, M = C.shape[1] - 1, C.shape[2] - 1
        g_scores = gC[:, :N, :M].to(ctx.in_dtype)
        g_alpha = gC[:, N, :].sum() + gC[:, :N, M].sum()          # every dustbin entry holds alpha (mdgat.py:294-299)
        return g_scores, g_alpha.reshape(()), None


def log_optimal_transport(scores, alpha, iters):
    """Differentiable Z (B,N+1,M+1) on the CUDA kernels; scores (B,N,M) float64, alpha a 0-dim tensor."""
    return LogOptimalTransportFn.apply(scores, alpha.reshape(()), int(iters))


def match_extract(C, u, v, loss_method='triplet_loss', mutual_check=False, match_threshold=0.2,
                  gt0=None, gt1=None, gamma=0.5, want_Z=False):
    """Match extraction (mdgat.py:442-483) and, with gt given, the loss of `loss_method` (mdgat.py:487-594) from the
    Sinkhorn outputs; 'loss' is a scalar, or (B,) for gap_loss. gt: -1 = no match (triplet / gap also accept M / N)."""
    _need_cuda(C)
    dev = C.device
    B, N, M = C.shape[0], C.shape[1] - 1, C.shape[2] - 1
    m0 = torch.empty((B, N), dtype=torch.int64, device=dev)
    m1 = torch.empty((B, M), dtype=torch.int64, device=dev)
    s0 = torch.empty((B, N), dtype=torch.float64, device=dev)
    s1 = torch.empty((B, M), dtype=torch.float64, device=dev)
    loss_mode = _capi.LOSS_NONE
    if gt0 is not None:
        loss_mode = {'triplet_loss': _capi.LOSS_TRIPLET, 'gap_loss': _capi.LOSS_GAP, 'superglue': _capi.LOSS_SUPERGLUE}[loss_method]
    loss = torch.zeros((B,) if loss_mode == _capi.LOSS_GAP else (), dtype=torch.float64, device=dev)
    nvalid = torch.zeros((), dtype=torch.int32, device=dev)
    Z = torch.empty_like(C) if want_Z else None
    scratch = torch.empty(_capi.lib.mdgat_match_scratch_doubles(B, N, M), dtype=torch.float64, device=dev)
    g0 = gt0.to(torch.int16).contiguous() if gt0 is not None else None
    g1 = gt1.to(torch.int16).contiguous() if gt1 is not None else None
    fout = _capi.ForwardOut(m0.data_ptr(), m1.data_ptr(), s0.data_ptr(), s1.data_ptr(), loss.data_ptr(),
                            nvalid.data_ptr(), Z.data_ptr() if Z is not None else None)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib.mdgat_match_extract(
            C.data_ptr(), u.data_ptr(), v.data_ptr(), B, N, M,
            _capi.MATCH_THRESHOLD if loss_method == 'superglue' else _capi.MATCH_DUSTBIN,
            int(bool(mutual_check)), float(match_threshold), loss_mode, float(gamma),
            g0.data_ptr() if g0 is not None else None, g1.data_ptr() if g1 is not None else None,
            ctypes.byref(fout), scratch.data_ptr(), _stream(dev)))
    return {'matches0': m0, 'matches1': m1, 'matching_scores0': s0, 'matching_scores1': s1,
            'loss': loss, 'nvalid0': nvalid, 'Z': Z}


def knn(x, src, k):
    """x (B,3,n), src (B,3,m) -> (B,n,k) int64 indices of the k nearest sources (mdgat.py:8-15)."""
    _need_cuda(x)
    x = x.double().contiguous()
    src = src.double().contiguous()
    B, _, n = x.shape
    m = src.shape[2]
    idx = torch.empty((B, n, k), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib.mdgat_knn(x.data_ptr(), src.data_ptr(), idx.data_ptr(), B, n, m, int(k), _stream(x.device)))
    return idx


def get_graph_feature(x, src, k):
    """One-hot kNN adjacency (B,n,m) int64 (mdgat.py:17-32)."""
    idx = knn(x, src, k)
    adj = torch.zeros((x.shape[0], x.shape[2], src.shape[2]), dtype=torch.int64, device=x.device)
    return adj.scatter_(2, idx, 1)


def encode(blob, data):
    """denc(desc) + kenc(kpts, scores) for both sides -> (desc0 (B,128,N), desc1 (B,128,M))."""
    k0 = data['keypoints0']
    _need_cuda(k0)
    dev = k0.device
    B, N, M = k0.shape[0], k0.shape[1], data['keypoints1'].shape[1]
    t = [data[k].double().contiguous() for k in ('keypoints0', 'keypoints1', 'descriptors0', 'descriptors1',
                                                 'scores0', 'scores1')]
    R = B * (N + M)
    X = torch.empty((R, LDX), dtype=torch.float64, device=dev)
    tmp = torch.empty(_capi.lib.mdgat_encode_scratch_doubles(R), dtype=torch.float64, device=dev)
    fin = _capi.ForwardIn(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(),
                          t[5].data_ptr(), None, None)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib.mdgat_encode(ctypes.byref(fin), B, N, M, _capi.F64, _capi.F64, blob.data_ptr(),
                                           X.data_ptr(), tmp.data_ptr(), _stream(dev)))
    d0 = X[:B * N, :128].reshape(B, N, 128).transpose(1, 2).contiguous()
    d1 = X[B * N:, :128].reshape(B, M, 128).transpose(1, 2).contiguous()
    return d0, d1


def measure_fp64_peak():
    a, b = ctypes.c_double(), ctypes.c_double()
    _capi.check(_capi.lib.mdgat_measure_fp64_peak(ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def measure_i8_peak():
    """Issue-rate ceiling of tcgen05.mma kind::i8 in TOP/s (2 operations per multiply-add)."""
    a = ctypes.c_double()
    _capi.check(_capi.lib.mdgat_measure_i8_peak(ctypes.byref(a)))
    return a.value


def measure_fp64_mixed():
    """(DMMA TF/s in a DMMA+DFMA mix, DFMA TF/s in the mix, DMMA TF/s of a register-tiled 4x4 loop)."""
    a, b = ctypes.c_double(), (ctypes.c_double * 2)()
    _capi.check(_capi.lib.mdgat_measure_fp64_mixed(ctypes.byref(a), b))
    return a.value, b[0], b[1]


def register_pairs(kpts0, kpts1, matches0, gt_matches0=None, T_gt=None):
    """Batched Kabsch registration + match statistics from predicted matches (device side of
    utils_test.solve_icp / calculate_error2 and the TP/FP/TN/FN counting of the eval scripts).
    Returns T (B,4,4) and a dict of per-pair tensors."""
    _need_cuda(kpts0)
    dev = kpts0.device
    if kpts0.dtype not in (torch.float32, torch.float64) or kpts1.dtype != kpts0.dtype:
        kpts0, kpts1 = kpts0.double(), kpts1.double()
    kpts0, kpts1 = kpts0.contiguous(), kpts1.contiguous()
    B, N, M = kpts0.shape[0], kpts0.shape[1], kpts1.shape[1]
    m0 = matches0.to(torch.int64).contiguous()
    g0 = gt_matches0.to(torch.int16).contiguous() if gt_matches0 is not None else None
    tg = T_gt.double().contiguous() if T_gt is not None else None
    T = torch.empty((B, 4, 4), dtype=torch.float64, device=dev)
    st = torch.empty((B, 8), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib.mdgat_register_pairs(
            kpts0.data_ptr(), kpts1.data_ptr(), _capi.F64 if kpts0.dtype == torch.float64 else _capi.F32,
            m0.data_ptr(), g0.data_ptr() if g0 is not None else None, tg.data_ptr() if tg is not None else None,
            B, N, M, T.data_ptr(), st.data_ptr(), _stream(dev)))
    names = ('n_valid', 'n_valid_gt', 'tp', 'fp', 'tn', 'fn', 'rte', 'rre')
    return T, {n: st[:, i] for i, n in enumerate(names)}


def prepare_pairs(kp1, kp2, pose1, pose2, T_cam0_velo, threshold, mutual_check=False):
    """Device version of the loader's per-item work (load_data.py:213-292): ground-truth matches,
    T_gt and the repeatability count for a batch of pairs. Returns (gt_matches0, gt_matches1, T_gt, rep)."""
    _need_cuda(kp1)
    dev = kp1.device
    kp1, kp2 = kp1.double().contiguous(), kp2.double().contiguous()
    pose1, pose2 = pose1.double().contiguous(), pose2.double().contiguous()
    calib = T_cam0_velo.double().contiguous()
    B, N, M = kp1.shape[0], kp1.shape[1], kp2.shape[1]
    m1 = torch.empty((B, N), dtype=torch.int16, device=dev)
    m2 = torch.empty((B, M), dtype=torch.int16, device=dev)
    T = torch.empty((B, 4, 4), dtype=torch.float64, device=dev)
    rep = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib.mdgat_prepare_pairs(kp1.data_ptr(), kp2.data_ptr(), pose1.data_ptr(), pose2.data_ptr(),
                                                  calib.data_ptr(), int(calib.dim() == 3), B, N, M, float(threshold),
                                                  int(bool(mutual_check)), m1.data_ptr(), m2.data_ptr(), T.data_ptr(),
                                                  rep.data_ptr(), _stream(dev)))
    return m1, m2, T, rep


def linear_i8(x, w, bias=None, relu=False, residual=None, x2=None, slices=7):
    """y = act([x | x2] w^T + bias) + residual through the tcgen05 int8 tensor cores (Ozaki splitting,
    float64-faithful). x (R,K0) [+ x2 (R,K1)], K in {128,256,512}; w (Nout,K) float64, Nout % 64 == 0."""
    from . import packing
    _need_cuda(x)
    x = x.double().contiguous()
    R, K0 = x.shape
    K1 = 0
    if x2 is not None:
        x2 = x2.double().contiguous()
        K1 = x2.shape[1]
    w = w.double().contiguous()
    nout = w.shape[0]
    wsl, cs = packing.slice_weight(w, slices)
    y = torch.empty((R, nout), dtype=torch.float64, device=x.device)
    b = bias.double().contiguous() if bias is not None else None
    r = residual.double().contiguous() if residual is not None else None
    scratch = torch.empty(_capi.lib.mdgat_linear_i8_scratch_bytes(R, K0 + K1, slices), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib.mdgat_linear_i8(
            x.data_ptr(), x.stride(0), K0, x2.data_ptr() if x2 is not None else None, x2.stride(0) if x2 is not None else 0, K1,
            wsl.data_ptr(), cs.data_ptr(), b.data_ptr() if b is not None else None,
            r.data_ptr() if r is not None else None, r.stride(0) if r is not None else 0,
            y.data_ptr(), y.stride(0), R, nout, int(relu), slices, scratch.data_ptr(), _stream(x.device)))
    return y
