"""Developer micro-benchmark: per-kernel CUDA-event times of the sequential (single-stream) full_demod."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

b200 = g.load_package()
b200.init(0)
shapes = [(256, 8192, 6), (1024, 8192, 6)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for n_ch, bp, nb in shapes:
    nsym = bp * nb // 10 + 1
    lv = torch.tensor([1.0, 3.0, -1.0, -3.0], device="cuda")[torch.randint(0, 4, (n_ch, nsym), device="cuda")]
    ph = torch.cumsum(lv.repeat_interleave(10, dim=1)[:, : bp * nb] * 0.0785, dim=1)
    iq = (torch.stack([0.85 * torch.cos(ph), 0.85 * torch.sin(ph)], dim=-1) + 0.05 * torch.randn((n_ch, bp * nb, 2), device="cuda")).contiguous().float()
    out = torch.empty((n_ch, bp * nb), device="cuda", dtype=torch.float32)
    bank = b200.DemodBank(n_ch, 48000, True)
    for _ in range(3):
        bank.full_demod(iq, bp, nb, out)
    torch.cuda.synchronize()
    b200.timing_enable(True)
    for _ in range(20):
        bank.full_demod(iq, bp, nb, out)
    torch.cuda.synchronize()
    rep = b200.timing_report()
    b200.timing_enable(False)
    print(f"ch={n_ch} bp={bp} nb={nb} debug={os.environ.get('DSDNEO_B200_REC_DEBUG', '0')}: " +
          "  ".join(f"{k}={v['ms'] / v['launches']:.4f}ms" if isinstance(v, dict) else f"{k}={v}" for k, v in rep.items()))
    bank.close()
