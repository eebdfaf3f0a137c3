"""Runs one cfg2-sized forward and dumps matches / scores (compare runs with MDGAT_FUSE_SLICE=0 and 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mdgat_matcher_b200 import synth
from mdgat_matcher_b200.models.mdgat import MDGAT
from oracle.ref_loader import net_config
dev = torch.device('cuda:0')
cfg = net_config(L=9, sinkhorn_iterations=100)
net = MDGAT(cfg); net.load_state_dict(synth.seeded_state_dict(9, 0)); net = net.double().eval().to(dev)
data = {k: v.to(dev) for k, v in synth.make_batch(3, 32, 512).items()}
out = net(data)
torch.cuda.synchronize()
np.savez(sys.argv[1], m0=out['matches0'].cpu().numpy(), m1=out['matches1'].cpu().numpy(), s0=out['matching_scores0'].cpu().numpy(), loss=out['loss'].cpu().numpy())
