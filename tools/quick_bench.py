"""Developer micro-benchmark (not the judged bench): times the block-side kernels with CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

b200 = g.load_package()
b200.init(0)


def time_it(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))


for n_ch, bp, nb in [(256, 8192, 6), (1024, 8192, 6), (4096, 8192, 2)]:
    for arith in (0, 1):
        iq = torch.randn((n_ch, bp * nb, 2), device="cuda", dtype=torch.float32) * 0.3
        out = torch.empty((n_ch, bp * nb), device="cuda", dtype=torch.float32)
        bank = b200.DemodBank(n_ch, 48000, True, fir_arith=arith)
        med, mn = time_it(lambda: bank.full_demod(iq, bp, nb, out))
        samples = n_ch * bp * nb
        print(f"full_demod ch={n_ch} bp={bp} nb={nb} arith={'fma' if arith == 0 else 'nofma'}: "
              f"median {med:.3f} ms  min {mn:.3f} ms  -> {samples / med / 1e6:.2f} GS/s, "
              f"{samples * 12 / med / 1e6:.1f} GB/s algorithmic, x{samples / med * 1e3 / (n_ch * 48000):.0f} real time")
        bank.close()
        del iq, out
