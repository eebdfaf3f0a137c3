"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference module as the parity pin.

Only tests/, tools that generate tests/golden/*, __graft_entry__.build() and bench.py's
cpu_baseline/--impl reference leg may import this file. The product path never does.

The reference hard-codes ``torch.device('cuda')`` inside its hot path
(/root/reference/models/mdgat.py:25,200,466-474,491-558), so on a CPU-only box the
unmodified file is run against a proxy object that stands in for the name ``torch`` inside
``models.mdgat`` only: every attribute forwards to real torch, but ``device(...)`` and
``device='cuda'`` keyword arguments resolve to the CPU (SURVEY.md section 8c recipe).

Weight convention = /root/reference/test.py:152-159,193: build the net in fp32, wrap in
DataParallel, ``load_state_dict`` (fp64 -> fp32 rounding), then ``.double().eval()``.
"""
import os
import sys
import types

import torch

def _reference_root():
    # /root/reference in the build container; on the GPU box the byte-identical copy oracle/build_ref.py placed under
    # oracle/_ref/reference (git-ignored build output that travels with the snapshot)
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.environ.get('MDGAT_REFERENCE_ROOT', '/root/reference')
    if not os.path.isfile(os.path.join(root, 'models', 'mdgat.py')):
        alt = os.path.join(here, '_ref', 'reference')
        if os.path.isfile(os.path.join(alt, 'models', 'mdgat.py')):
            return alt
    return root


REFERENCE_ROOT = _reference_root()
CHECKPOINT = os.path.join(REFERENCE_ROOT, 'pre-trained', 'best_model.pth')

DEFAULT_K = [128, None, 128, None, 64, None, 64, None]      # test.py:83


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'models', 'mdgat.py'))


class _TorchOnCpu:
    """Stands in for the module-global name ``torch`` inside the reference file."""

    def __init__(self, target):
        self._target = torch.device(target)

    def device(self, *a, **k):
        return self._target

    def _remap(self, kwargs):
        dev = kwargs.get('device', None)
        if dev is not None and 'cuda' in str(dev):
            kwargs['device'] = self._target
        return kwargs

    def zeros_like(self, *a, **k):
        return torch.zeros_like(*a, **self._remap(k))

    def zeros(self, *a, **k):
        return torch.zeros(*a, **self._remap(k))

    def arange(self, *a, **k):
        return torch.arange(*a, **self._remap(k))

    def __getattr__(self, name):
        return getattr(torch, name)


def load_reference_module(name='mdgat', target='cpu'):
    """Import /root/reference/models/<name>.py unmodified; returns the module object."""
    if not reference_available():
        raise FileNotFoundError('reference tree not present at %s' % REFERENCE_ROOT)
    sys.dont_write_bytecode = True            # reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    saved = {k: sys.modules.get(k) for k in ('models', 'models.mdgat', 'models.superglue')}
    for k in saved:
        sys.modules.pop(k, None)
    try:
        mod = importlib.import_module('models.' + name)
    finally:
        # do not leave the reference registered under 'models.*' (the drop-in uses that name)
        ref_mods = {k: sys.modules.pop(k, None) for k in list(sys.modules)
                    if k == 'models' or k.startswith('models.')}
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
        _KEEP.update(ref_mods)
    if str(target) == 'cpu' or not torch.cuda.is_available():
        mod.torch = _TorchOnCpu('cpu')
    return mod


_KEEP = {}


def net_config(L=9, k=None, sinkhorn_iterations=100, loss_method='triplet_loss',
               mutual_check=False, descriptor='FPFH', match_threshold=0.2):
    """The dict test.py:137-151 builds."""
    return {
        'sinkhorn_iterations': sinkhorn_iterations,
        'match_threshold': match_threshold,
        'lr': 1e-4,
        'loss_method': loss_method,
        'k': list(DEFAULT_K) if k is None else k,
        'descriptor': descriptor,
        'mutual_check': mutual_check,
        'triplet_loss_gamma': 0.5,
        'train_step': 3,
        'L': L,
    }


def build_reference_net(cfg, weights='checkpoint', seed=0, target='cpu'):
    """Returns (net, module, zcap). net is DataParallel(MDGAT) in fp64 eval mode.

    weights: 'checkpoint' (needs L=9) | 'seeded' (torch.manual_seed(seed) init + the
    make_matchy() adjustment so that real matches appear) | a state_dict.
    zcap['Z'] holds the last assignment matrix (captured by wrapping log_optimal_transport).
    """
    mod = load_reference_module('mdgat', target)
    if isinstance(weights, str) and weights == 'seeded':
        torch.manual_seed(seed)
    net = mod.MDGAT(cfg)                                  # fp32 params (test.py:156)
    net = torch.nn.DataParallel(net)                      # test.py:158
    if isinstance(weights, str) and weights == 'checkpoint':
        if os.path.isfile(CHECKPOINT):
            ck = torch.load(CHECKPOINT, map_location='cpu', weights_only=True)
            net.load_state_dict(ck['net'])                # fp64 -> fp32 copy (test.py:159)
        else:
            # GPU box: the 71 MB checkpoint did not travel; oracle/_ref/best_model_fp32.npz holds exactly the fp32
            # values that copy produces
            from oracle.build_ref import load_checkpoint_state_dict
            sd = load_checkpoint_state_dict()
            if sd is None:
                raise FileNotFoundError('neither %s nor oracle/_ref/best_model_fp32.npz exists' % CHECKPOINT)
            net.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in sd.items()})
    elif isinstance(weights, str) and weights == 'seeded':
        pass                                              # torch.manual_seed(seed) default initialisation
    else:
        net.load_state_dict(weights)
    if str(target) != 'cpu':
        net.to(torch.device(target))                      # test.py:172
    else:
        net.module.to('cpu')                              # DataParallel.__init__ moves the module to cuda:0 when exactly one GPU is visible
    net.double().eval()                                   # test.py:193
    zcap = {}
    orig = mod.log_optimal_transport

    def _capture(scores, alpha, iters):
        zcap['scores_in'] = scores.detach().clone()
        z = orig(scores, alpha, iters)
        zcap['Z'] = z.detach().clone()
        return z

    mod.log_optimal_transport = _capture
    return net, mod, zcap


def run_reference(net, data):
    """forward() mutates gt_matches in place (mdgat.py:519-520); hand it copies."""
    d = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in data.items()}
    # With a GPU visible, DataParallel.forward scatters its inputs to cuda:0 even when the wrapped module lives on the
    # CPU; the CPU oracle / baseline therefore calls the wrapped (unmodified) MDGAT module directly.
    call = net
    if isinstance(net, torch.nn.DataParallel) and next(net.parameters()).device.type == 'cpu':
        call = net.module
    with torch.no_grad():
        out = call(d)
    return out
