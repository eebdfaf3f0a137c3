"""Time and peak memory of one training step (forward + backward of MDGAT in train mode, the torch path of
models/mdgat.py::_forward_torch) with the Sinkhorn stage on the CUDA kernels (fused forward + hand-written backward,
csrc/sinkhorn_bwd.cu) against autograd through the unrolled torch iterations. Development tool.

    python tools/train_step_probe.py [batch] [n_keypoints] [sinkhorn_iters]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import net_config, load_weights     # noqa: E402


def main():
    from mdgat_matcher_b200 import synth
    from mdgat_matcher_b200.models.mdgat import MDGAT
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    dev = torch.device('cuda:0')
    sd, _ = load_weights(9)
    data = {k: v.to(dev) for k, v in synth.make_batch(5, B, n).items()}
    res = {'batch': B, 'n': n, 'sinkhorn_iterations': T}
    for name, cuda_attn, cuda_sk in (('cuda_attention+sinkhorn', True, True), ('cuda_sinkhorn', False, True), ('torch', False, False)):
        cfg = net_config(9, T)
        cfg['cuda_sinkhorn_backward'] = cuda_sk
        cfg['cuda_attention_backward'] = cuda_attn
        net = MDGAT(cfg)
        net.load_state_dict(sd)
        net = net.double().train().to(dev)
        times = []
        for it in range(4):
            net.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats(dev)
            t0 = time.perf_counter()
            d = dict(data)
            d['gt_matches0'], d['gt_matches1'] = data['gt_matches0'].clone(), data['gt_matches1'].clone()
            out = net(d)
            out['loss'].backward()
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        res[name] = {
            'step_ms': round(min(times[1:]), 2), 'peak_gb': round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 3),
            'loss': float(out['loss'].detach())}
        del net, out
        torch.cuda.empty_cache()
    print(json.dumps(res))


if __name__ == '__main__':
    main()
