import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
b200 = g.load_package(); b200.init(0)
n_ch, bp, nb = 1024, 8192, 2
iq = torch.randn((n_ch, bp * nb, 2), device="cuda") * 0.3
out = torch.empty((n_ch, bp * nb), device="cuda")
bank = b200.DemodBank(n_ch, 48000, True)
for _ in range(3):
    bank.full_demod(iq, bp, nb, out)
torch.cuda.synchronize()
