"""ctypes binding of libmdgat_b200.so (include/mdgat_b200.h). No torch types cross this
boundary: callers pass raw device pointers (tensor.data_ptr()) and a cudaStream_t handle.

There is no fallback: if the library is missing the import fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libmdgat_b200.so')

MDGAT_OK = 0
MATCH_DUSTBIN, MATCH_THRESHOLD = 0, 1
LOSS_NONE, LOSS_TRIPLET, LOSS_GAP, LOSS_SUPERGLUE = 0, 1, 2, 3
F32, F64 = 0, 1
GEMM_DMMA_F64, GEMM_TCGEN05_I8 = 0, 1
ATTN_DMMA_F64, ATTN_TCGEN05_I8, ATTN_TCGEN05_I8_ALL = 0, 1, 2
LDX = 132
LDH_QK, LDH_V = 36, 34


class ForwardCfg(C.Structure):
    _fields_ = [('B', C.c_int), ('N', C.c_int), ('M', C.c_int), ('L', C.c_int),
                ('sinkhorn_iters', C.c_int), ('layer_k', C.POINTER(C.c_int)),
                ('match_mode', C.c_int), ('mutual_check', C.c_int), ('match_threshold', C.c_double),
                ('loss_mode', C.c_int), ('triplet_gamma', C.c_double),
                ('in_dtype', C.c_int), ('score_dtype', C.c_int), ('write_Z', C.c_int),
                ('gemm_mode', C.c_int), ('gemm_slices', C.c_int), ('attn_mode', C.c_int),
                ('attn_slices', C.c_int), ('attn_p_slices', C.c_int), ('sinkhorn_k32', C.c_int),
                ('late_from', C.c_int), ('late_gemm_slices', C.c_int), ('late_attn_slices', C.c_int),
                ('late_attn_p_slices', C.c_int), ('d_weights_i8_late', C.c_void_p)]


class ForwardIn(C.Structure):
    _fields_ = [('d_kpts0', C.c_void_p), ('d_kpts1', C.c_void_p), ('d_desc0', C.c_void_p),
                ('d_desc1', C.c_void_p), ('d_scores0', C.c_void_p), ('d_scores1', C.c_void_p),
                ('d_gt0', C.c_void_p), ('d_gt1', C.c_void_p)]


class ForwardOut(C.Structure):
    _fields_ = [('d_matches0', C.c_void_p), ('d_matches1', C.c_void_p), ('d_mscores0', C.c_void_p),
                ('d_mscores1', C.c_void_p), ('d_loss', C.c_void_p), ('d_nvalid0', C.c_void_p),
                ('d_Z', C.c_void_p)]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            'mdgat-matcher_b200: %s is missing. Build it with `python __graft_entry__.py build` '
            '(nvcc, sm_100a). There is no CPU or eager fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i, d, ll, sz = C.c_void_p, C.c_int, C.c_double, C.c_longlong, C.c_size_t
    sig = {
        'mdgat_last_error': (C.c_char_p, []),
        'mdgat_abi_version': (i, []),
        'mdgat_weight_blob_doubles': (sz, [i]),
        'mdgat_forward_workspace_bytes': (sz, [C.POINTER(ForwardCfg)]),
        'mdgat_forward': (i, [C.POINTER(ForwardCfg), vp, vp, C.POINTER(ForwardIn), C.POINTER(ForwardOut), vp, sz, vp]),
        'mdgat_linear_i8_scratch_bytes': (sz, [i, i, i]),
        'mdgat_linear_i8': (i, [vp, i, i, vp, i, i, vp, vp, vp, vp, i, vp, i, i, i, i, i, vp, vp]),
        'mdgat_linear_f64': (i, [vp, i, i, vp, i, i, vp, i, vp, vp, i, vp, i, i, i, d, i, vp]),
        'mdgat_gemm_nt_f64': (i, [vp, i, ll, vp, i, ll, vp, i, ll, i, i, i, i, d, vp]),
        'mdgat_encode_scratch_doubles': (sz, [i]),
        'mdgat_encode': (i, [C.POINTER(ForwardIn), i, i, i, i, i, vp, vp, vp, vp]),
        'mdgat_attention_f64_scratch_doubles': (sz, [i, i, i]),
        'mdgat_attention_f64': (i, [vp, vp, vp, vp, i, i, i, i, i, vp, vp]),
        'mdgat_attention_i8_scratch_bytes': (sz, [i, i, i]),
        'mdgat_attention_i8': (i, [vp, vp, vp, vp, i, i, i, i, i, vp, vp, i, i, vp]),
        'mdgat_sinkhorn_scratch_doubles': (sz, [i, i, i]),
        'mdgat_sinkhorn_read_status': (i, [vp, i, i, i, C.POINTER(i), C.POINTER(i)]),
        'mdgat_forward_sinkhorn_status': (i, [C.POINTER(ForwardCfg), vp, C.POINTER(i), C.POINTER(i)]),
        'mdgat_sinkhorn_f64': (i, [vp, vp, vp, vp, i, i, i, i, vp, vp]),
        'mdgat_attention_backward_scratch_doubles': (sz, [i, i, i, i]),
        'mdgat_attention_backward_f64': (i, [vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, vp, vp]),
        'mdgat_sinkhorn_backward_scratch_doubles': (sz, [i, i, i, i]),
        'mdgat_sinkhorn_backward_f64': (i, [vp, vp, vp, i, i, i, i, vp, C.POINTER(i), vp]),
        'mdgat_sinkhorn_f64_k32': (i, [vp, vp, vp, vp, i, i, i, i, vp, vp]),
        'mdgat_match_scratch_doubles': (sz, [i, i, i]),
        'mdgat_match_extract': (i, [vp, vp, vp, i, i, i, i, i, d, i, d, vp, vp, C.POINTER(ForwardOut), vp, vp]),
        'mdgat_knn': (i, [vp, vp, vp, i, i, i, i, vp]),
        'mdgat_prepare_pairs': (i, [vp, vp, vp, vp, vp, i, i, i, i, d, i, vp, vp, vp, vp, vp]),
        'mdgat_register_pairs': (i, [vp, vp, i, vp, vp, vp, i, i, i, vp, vp, vp]),
        'mdgat_measure_fp64_peak': (i, [C.POINTER(d), C.POINTER(d)]),
        'mdgat_measure_fp64_mixed': (i, [C.POINTER(d), C.POINTER(d)]),
        'mdgat_measure_i8_peak': (i, [C.POINTER(d)]),
        'mdgat_launch_count': (ll, []),
        'mdgat_launch_count_add': (None, [ll]),
        'mdgat_debug_trace': (i, [vp]),
        'mdgat_debug_flags': (i, [i]),
        'mdgat_profile_enable': (i, [i]),
        'mdgat_profile_collect': (i, [C.POINTER(d), C.POINTER(ll), C.POINTER(ll), i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTS = _load()


STAGES = ('encode', 'gemm', 'attn_full', 'attn_topk', 'sinkhorn', 'match', 'slice')


def profile_collect():
    n = len(STAGES)
    ms, la, sg = (C.c_double * n)(), (C.c_longlong * n)(), (C.c_longlong * n)()
    check(lib.mdgat_profile_collect(ms, la, sg, n))
    return {s: {'ms': ms[i], 'launches': la[i], 'segments': sg[i]} for i, s in enumerate(STAGES)}


class MdgatError(RuntimeError):
    pass


def check(code):
    if code != MDGAT_OK:
        raise MdgatError(lib.mdgat_last_error().decode())
