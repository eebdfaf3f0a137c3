"""Runs the three per-layer Ozaki GEMM shapes of cfg2 (R = 32768 rows) a few times -- the target of ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdgat_matcher_b200 import ops
R = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device('cuda:0')
g = torch.Generator(device='cpu').manual_seed(0)
x = torch.randn(R, 128, generator=g, dtype=torch.float64).to(dev)
m = torch.randn(R, 128, generator=g, dtype=torch.float64).to(dev)
h = torch.randn(R, 256, generator=g, dtype=torch.float64).to(dev)
wqkv = torch.randn(384, 128, generator=g, dtype=torch.float64).to(dev)
w1 = torch.randn(256, 256, generator=g, dtype=torch.float64).to(dev)
w2 = torch.randn(128, 256, generator=g, dtype=torch.float64).to(dev)
b1 = torch.randn(256, generator=g, dtype=torch.float64).to(dev)
for it in range(3):
    ops.linear_i8(x, wqkv)
    ops.linear_i8(x, w1, bias=b1, relu=True, x2=m)
    ops.linear_i8(h, w2, residual=x)
torch.cuda.synchronize()
