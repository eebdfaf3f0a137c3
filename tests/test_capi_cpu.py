"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/mdgat_b200.h declares; the weight packer and the k schedule; the drop-in module keeps
the reference's state-dict layout. No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from mdgat_matcher_b200 import _capi
    hdr = open(os.path.join(ROOT, 'include', 'mdgat_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(mdgat_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_capi.EXPORTS) == declared
    assert _capi.lib.mdgat_abi_version() == 4


def test_error_text_without_gpu():
    from mdgat_matcher_b200 import _capi
    rc = _capi.lib.mdgat_knn(None, None, None, 1, 4, 4, 9, None)          # k > m is rejected before any CUDA call
    assert rc == -1 and b'out of range' in _capi.lib.mdgat_last_error()


def test_blob_size_and_k_schedule():
    from mdgat_matcher_b200 import _capi, packing, synth
    for L in (1, 4, 9):
        assert packing.blob_doubles(L) == _capi.lib.mdgat_weight_blob_doubles(L)
        blob = packing.pack_state_dict(synth.seeded_state_dict(L, 1), L)
        assert blob.dtype == torch.float64 and blob.numel() == packing.blob_doubles(L)
    k = [128, None, 128, None, 64, None, 64, None]
    assert packing.layer_k_schedule(k, 9) == [0] * 10 + [128, 0, 128, 0, 64, 0, 64, 0]
    assert packing.layer_k_schedule(k, 4) == [128, 0, 128, 0, 64, 0, 64, 0]
    assert packing.layer_k_schedule([], 9) == [0] * 18
    from oracle import mdgat_oracle as O
    for L in (4, 9):
        assert [O.layer_topk(i, k, L) or 0 for i in range(2 * L)] == packing.layer_k_schedule(k, L)


def test_packed_blob_reproduces_reference_layers_on_cpu():
    """Folded-BN + head permutation: emulate one GNN layer from the blob with numpy and compare
    with the oracle's unfused layer."""
    from mdgat_matcher_b200 import packing, synth
    from oracle import mdgat_oracle as O
    L = 1
    sd_t = synth.seeded_state_dict(L, 3)
    sd = O.state_dict_to_numpy(sd_t)
    blob = packing.pack_state_dict(sd_t, L).numpy()
    off = sum(packing.tiled_doubles(packing.KENC_DIMS[i + 1], packing.KENC_DIMS[i]) + packing.KENC_DIMS[i + 1] for i in range(4))
    off += sum(packing.tiled_doubles(packing.DENC_DIMS[i + 1], packing.DENC_DIMS[i]) + packing.DENC_DIMS[i + 1] for i in range(3))

    def take(n, shape=None):
        nonlocal off
        a = blob[off:off + n]
        off += n
        return a.reshape(shape) if shape else a

    def take_w(nout, k):          # undo packing.tile_weight
        t = take(packing.tiled_doubles(nout, k)).reshape(-(-nout // 128), -(-k // 32), 128, 36)[..., :32]
        return t.transpose(0, 2, 1, 3).reshape(t.shape[0] * 128, t.shape[1] * 32)[:nout, :k]
    wqkv, bqkv = take_w(384, 128), take(384)
    w1, b1 = take_w(256, 256), take(256)
    w2, b2 = take_w(128, 256), take(128)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(1, 128, 40)); src = rng.normal(size=(1, 128, 50))
    want, _ = O.attentional_propagation(sd, 'gnn.layers.0', x, src, None)
    xr, sr = x[0].T, src[0].T                                  # point-major rows
    q = xr @ wqkv[:128].T + bqkv[:128]; k = sr @ wqkv[128:256].T + bqkv[128:256]; v = sr @ wqkv[256:].T + bqkv[256:]
    msg = np.zeros((40, 128))
    for h in range(4):
        s = q[:, h * 32:(h + 1) * 32] @ k[:, h * 32:(h + 1) * 32].T / np.sqrt(32)
        p = np.exp(s - s.max(1, keepdims=True)); p /= p.sum(1, keepdims=True)
        msg[:, h * 32:(h + 1) * 32] = p @ v[:, h * 32:(h + 1) * 32]
    hd = np.maximum(np.concatenate([xr, msg], 1) @ w1.T + b1, 0)      # merge conv is folded into w1 / b1
    delta = hd @ w2.T + b2
    assert np.abs(delta.T - want[0]).max() < 1e-11


def test_dropin_module_state_dict_layout_and_checkpoint():
    from mdgat_matcher_b200.models.mdgat import MDGAT
    from mdgat_matcher_b200.models.superglue import SuperGlue
    from mdgat_matcher_b200 import synth
    from conftest import case_cfg
    cfg = case_cfg({'L': 9, 'T': 100})
    net = MDGAT(cfg)
    keys = set(net.state_dict().keys())
    assert len(keys) == 348 and keys == set(synth.seeded_state_dict(9).keys())
    for k, v in synth.seeded_state_dict(9).items():
        assert tuple(net.state_dict()[k].shape) == tuple(v.shape), k
    from oracle.build_ref import load_checkpoint_state_dict
    sd = load_checkpoint_state_dict()
    if sd is not None:
        wrapped = torch.nn.DataParallel(net)                 # test.py:158-159 loads through DataParallel
        wrapped.load_state_dict({'module.' + k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        SuperGlue(cfg).load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    with pytest.raises(KeyError):
        MDGAT({k: v for k, v in cfg.items() if k != 'L'})
    with pytest.raises(Exception, match='Invalid descriptor'):
        MDGAT({**cfg, 'descriptor': 'nope'})


def test_eval_forward_refuses_cpu_tensors_and_train_path_matches_oracle():
    from mdgat_matcher_b200.models.mdgat import MDGAT
    from mdgat_matcher_b200 import synth
    from oracle import mdgat_oracle as O
    from conftest import case_cfg
    cfg = case_cfg({'L': 2, 'T': 10, 'k': [16, None]})
    sd = synth.seeded_state_dict(2, 5)
    net = MDGAT(cfg)
    net.load_state_dict(sd)
    net.double().eval()
    data = synth.make_batch(2, 2, 48)
    with pytest.raises(RuntimeError, match='CUDA'):
        net({k: v.clone() for k, v in data.items()})
    with torch.no_grad():
        out = net._forward_torch({k: v.clone() for k, v in data.items()})      # differentiable path, eval BN
    want = O.forward(O.state_dict_to_numpy(sd), data, cfg)
    assert np.array_equal(out['matches0'].numpy(), want['matches0'])
    assert np.abs(out['matching_scores0'].numpy() - want['matching_scores0']).max() < 1e-9
    assert abs(float(out['loss']) - float(want['loss'])) < 1e-9
    # train mode is differentiable end to end
    net.train()
    out = net({k: v.clone() for k, v in data.items()})
    out['loss'].backward()
    assert net.final_proj.weight.grad is not None and torch.isfinite(net.final_proj.weight.grad).all()


def test_slice_weight_digit_planes_are_exact_and_in_layout():
    """packing.slice_weight (operand of the tcgen05 int8 GEMM, csrc/ozaki_gemm.cu): digits |g| <= 64, 2^f * sum_s g_s 2^(1-7s)
    reproduces every weight to 2^-49 of its row-chunk maximum, colscale is an exact power of two, and the bytes sit in the
    UMMA canonical no-swizzle K-major order [col tile][k chunk][slice][(r/8)*1024 + (k/16)*128 + (r%8)*16 + k%16]."""
    from mdgat_matcher_b200 import packing
    g = torch.Generator().manual_seed(3)
    nout, k, S = 64, 256, 7
    w = torch.randn(nout, k, generator=g, dtype=torch.float64) * torch.exp(torch.randn(nout, 1, generator=g, dtype=torch.float64) * 3)
    w[5] = 0.0                                                     # an all-zero row keeps scale 1 and zero digits
    w[:, ::9] *= 1e-7
    planes, cs = packing.slice_weight(w, S)
    assert planes.dtype == torch.int8 and planes.numel() == (nout // 32) * (k // 128) * S * 32 * 128
    assert cs.shape == (k // 128, nout)
    m, e = torch.frexp(cs)
    assert torch.all(m == 0.5), 'colscale must be a power of two'
    d = planes.reshape(nout // 32, k // 128, S, 4, 8, 8, 16).to(torch.float64)      # [ct][kc][s][r/8][k/16][r%8][k%16]
    assert float(d.abs().max()) <= 64
    d = d.permute(0, 3, 5, 1, 4, 6, 2).reshape(nout, k, S)                            # [row][k][s]
    weights = torch.tensor([2.0 ** (1 - 7 * (s + 1)) for s in range(S)], dtype=torch.float64)
    scale = cs.t().repeat_interleave(128, dim=1)                                     # [row][k]
    rec = (d * weights).sum(-1) * scale
    rowmax = w.reshape(nout, k // 128, 128).abs().amax(-1).repeat_interleave(128, dim=1)
    err = (rec - w).abs()
    assert torch.all(err <= rowmax * 2.0 ** -48 + 0.0), float((err / rowmax.clamp_min(1e-300)).max())
    assert float(rec[5].abs().max()) == 0.0


def test_i8_blob_layout_matches_the_forward_offsets():
    """pack_state_dict_i8: per layer 36 slice tiles (12 q/k/v + 16 MLP conv 0 + 8 MLP conv 3, k chunks included) followed by
    1152 column scales -- the offsets mdgat_forward() hard-codes (capi.cu)."""
    from mdgat_matcher_b200 import packing, synth
    L, S = 2, 7
    blob = packing.pack_state_dict_i8(synth.seeded_state_dict(L, 0), L, S)
    per_layer = S * 32 * 128 * 36 + 1152 * 8
    assert blob.dtype == torch.uint8 and blob.numel() == 2 * L * per_layer == 2 * L * packing.i8_layer_bytes(S)
    cs = blob[S * 32 * 128 * 36: per_layer].view(torch.float64)
    m, _ = torch.frexp(cs)
    assert torch.all(m == 0.5)


def test_packed_weight_cache_invalidation_and_dataparallel_replicas():
    """The packed-weight cache: keyed on (storage, version, dtype), explicit invalidate_packed() for `.data` edits,
    dropped by load_state_dict(); DataParallel replicas share the source module's cache by reference and never pack
    themselves (their parameters are per-forward broadcast copies, not even registered when autograd is on)."""
    from mdgat_matcher_b200.models.mdgat import MDGAT
    from mdgat_matcher_b200 import synth
    from conftest import case_cfg
    cfg = case_cfg({'L': 2, 'T': 10, 'k': [16, None]})
    sd = synth.seeded_state_dict(2, 5)
    net = MDGAT(cfg)
    net.load_state_dict(sd)
    net.double().eval()
    b0 = net.packed_weights()
    assert b0.device.type == 'cpu' and net.packed_weights() is b0
    net.double()                                              # test.py:193 every iteration: no repack
    assert net.packed_weights() is b0
    bin_at = b0.numel() - 4
    # writes through .data bypass the version counter: documented, remedied by invalidate_packed()
    net.bin_score.data.fill_(3.0)
    assert net.packed_weights() is b0
    net.invalidate_packed()
    b1 = net.packed_weights()
    assert b1 is not b0 and float(b1[bin_at]) == 3.0
    with torch.no_grad():
        net.bin_score.add_(1.0)                               # ordinary in-place update: seen by the key
    assert float(net.packed_weights()[bin_at]) == 4.0
    net.load_state_dict(sd)                                   # always invalidates
    assert float(net.packed_weights()[bin_at]) == float(sd['bin_score'])
    i8 = net.packed_weights_i8(7)
    assert i8.dtype == torch.uint8 and net.packed_weights_i8(7) is i8
    # what torch.nn.parallel.replicate() does with autograd enabled: a shallow copy without registered parameters
    rep = net._replicate_for_data_parallel()
    rep._parameters.clear()
    assert rep._is_replica and not net._is_replica and 'bin_score' not in rep.state_dict()
    assert rep.packed_weights(torch.device('cpu')) is net.packed_weights()
    assert rep.packed_weights_i8(7, torch.device('cpu')) is i8
    assert rep._workspaces is net._workspaces


def test_forward_routing_unknown_loss_and_eval_autograd():
    from mdgat_matcher_b200.models.mdgat import MDGAT
    from mdgat_matcher_b200 import synth
    from conftest import case_cfg
    cfg = case_cfg({'L': 2, 'T': 10, 'k': [16, None]})
    sd = synth.seeded_state_dict(2, 5)
    data = synth.make_batch(2, 2, 48)
    bad = MDGAT({**cfg, 'loss_method': 'nope'})
    with pytest.raises(UnboundLocalError):                    # what falling off mdgat.py:486-603 raises
        bad.double().eval()({k: v.clone() for k, v in data.items()})
    # eval-mode call that must be differentiable (frozen-BatchNorm fine-tuning): opt-in torch path, CPU tensors fine
    net = MDGAT({**cfg, 'eval_autograd': True})
    net.load_state_dict(sd)
    net.double().eval()
    out = net({k: v.clone() for k, v in data.items()})
    assert out['loss'].requires_grad
    out['loss'].backward()
    assert net.final_proj.weight.grad is not None and float(net.final_proj.weight.grad.abs().sum()) > 0
    with torch.no_grad(), pytest.raises(RuntimeError, match='CUDA'):
        net({k: v.clone() for k, v in data.items()})          # autograd off: the CUDA path, which refuses CPU tensors
