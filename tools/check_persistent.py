"""Dumps the outputs of one large forward (many row tiles per CTA in the persistent Ozaki GEMM); run once with
MDGAT_OZ_PERSISTENT=0 and once with 1 and compare the files: the two schedules must agree bit for bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mdgat_matcher_b200 import synth
from mdgat_matcher_b200.models.mdgat import MDGAT
import bench
B, N = int(sys.argv[2]), int(sys.argv[3])
dev = torch.device('cuda:0')
cfg = bench.net_config(9, 20)
net = MDGAT(cfg); net.load_state_dict(synth.seeded_state_dict(9, 0)); net = net.double().eval().to(dev)
out = net({k: v.to(dev) for k, v in synth.make_batch(5, B, N).items()})
torch.cuda.synchronize()
np.savez(sys.argv[1], **{k: out[k].cpu().numpy() for k in ('matches0', 'matches1', 'matching_scores0', 'matching_scores1', 'loss')})
