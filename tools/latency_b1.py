"""Batch-1 latency of MDGAT.forward (cfg5 / cfg1 shapes) per engine set, with and without the CUDA graph, next to the
unmodified reference as eager PyTorch on the same GPU. Development tool (the chosen defaults are reported by bench.py).

    python tools/latency_b1.py [n_keypoints ...]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import net_config, load_weights     # noqa: E402


def timed(call, n=40):
    for _ in range(5):
        o = call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        o = call()
        float(o['loss'])
        ts.append((time.perf_counter() - t0) * 1e3)
    return {'median_ms': round(float(np.median(ts)), 4), 'min_ms': round(float(min(ts)), 4)}


def main():
    from mdgat_matcher_b200 import synth
    from mdgat_matcher_b200.models.mdgat import MDGAT
    dev = torch.device('cuda:0')
    L = 9
    sd, _ = load_weights(L)
    for n in [int(a) for a in sys.argv[1:]] or [256, 128, 512]:
        for B in (1, 4):
            data = {k: v.to(dev) for k, v in synth.make_batch(77, B, n).items()}
            res = {'n': n, 'B': B}
            for gemm, attn in (('tcgen05_i8', 'tcgen05_i8'), ('dmma', 'dmma'), ('dmma', 'tcgen05_i8'), ('tcgen05_i8', 'dmma')):
                for graph in (False, True):
                    cfg = net_config(L, 20)
                    cfg.update({'cuda_graph': graph, 'gemm': gemm, 'attention': attn})
                    net = MDGAT(cfg)
                    net.load_state_dict(sd)
                    net = net.double().eval().to(dev)
                    with torch.no_grad():
                        def call():
                            d = dict(data)
                            d['gt_matches0'], d['gt_matches1'] = data['gt_matches0'].clone(), data['gt_matches1'].clone()
                            return net(d)
                        res['%s+%s%s' % (gemm, attn, '+graph' if graph else '')] = timed(call)
            try:
                from oracle import ref_loader as RL
                rnet, mod, _z = RL.build_reference_net(RL.net_config(L=L, sinkhorn_iterations=20), 'checkpoint', target='cuda:0')
                with torch.no_grad():
                    def rcall():
                        d = {k: v.clone() for k, v in data.items()}
                        return rnet.module(d)
                    res['reference_eager'] = timed(rcall, 15)
            except Exception as e:
                res['reference_eager'] = repr(e)[:200]
            print(json.dumps(res), flush=True)


if __name__ == '__main__':
    main()
