"""One launch of the general channelizer per configuration (for ncu): python tools/pfbn_one.py M R [n_out]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import __graft_entry__ as G

b200 = G.load_package()
b200.init(0)
M, R = int(sys.argv[1]), int(sys.argv[2])
n_out = int(sys.argv[3]) if len(sys.argv) > 3 else 24576
x = torch.randint(0, 256, (n_out * M, 2), dtype=torch.uint8, device="cuda:0")
cz = b200.Channelizer(M, 8, True)
out = torch.empty((M // R, n_out, 2), device="cuda:0")
for _ in range(3):
    cz.channelize_bins(x, R, R - 1, out)
torch.cuda.synchronize()
